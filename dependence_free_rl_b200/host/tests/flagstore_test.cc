// flagstore_test -- xeno::flagstore (host/xeno/configuration.h) and the run-time problem definition of
// apps/bin_packing, without touching the device (runs in the CPU suite).
#include <cstdio>
#include <string>
#include <vector>

#include <xeno/configuration.h>

#include <apps/bin_packing/bin_packing.h>

static int fails = 0;
#define EXPECT(c)                                             \
  do {                                                        \
    if (!(c)) {                                               \
      std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); \
      ++fails;                                                \
    }                                                         \
  } while (0)

int main() {
  xeno::flagstore f;
  f.define_flag<int>("num_bins", 'b', 8);
  f.define_flag<double>("rate", 'r', 1e-4);
  f.define_flag<std::string>("nets", 'n', "ref");
  f.define_flag<bool>("verbose", 'v', false);
  EXPECT(f.get_flag<int>("num_bins") == 8 && f.get_flag<int>('b') == 8);
  std::vector<std::string_view> argv = {"--num_bins=32", "4096", "-r", "0.5", "--nets", "c2", "-v", "10", "--", "--not-a-flag"};
  std::vector<std::string_view> pos = f.parse_from_args(argv);
  EXPECT(f.get_flag<int>("num_bins") == 32);
  EXPECT(f.get_flag<double>("rate") == 0.5);
  EXPECT(f.get_flag<std::string>("nets") == "c2");
  EXPECT(f.get_flag<bool>("verbose"));
  EXPECT(pos.size() == 3 && pos[0] == "4096" && pos[1] == "10" && pos[2] == "--not-a-flag");
  f.set_flag("num_bins", 5);
  EXPECT(f.get_flag<int>('b') == 5);
  bool threw = false;
  try {
    f.get_flag<double>("num_bins");  // wrong type
  } catch (const xeno::error &) {
    threw = true;
  }
  EXPECT(threw);
  threw = false;
  try {
    std::vector<std::string_view> bad = {"--undefined=1"};
    f.parse_from_args(bad);
  } catch (const xeno::error &) {
    threw = true;
  }
  EXPECT(threw);
  threw = false;
  try {
    std::vector<std::string_view> bad = {"--num_bins=seven"};
    f.parse_from_args(bad);
  } catch (const xeno::error &) {
    threw = true;
  }
  EXPECT(threw);
  // the problem definition as launch parameters: bp::num_bins reads like the reference's constant
  bp::problem p;
  p.num_bins = 16;
  p.capacity = {16, 16};
  bp::configure(p);
  EXPECT(4 * bp::num_bins == 64 && bp::observation::length() == 64 && bp::action::cardinality() == 16);
  bp::observation ob(p.capacity);
  EXPECT(ob.bins.size() == bp::num_bins);
  ob.item = {4, 2};
  xylo::vector x = xylo::to_vector(ob);
  EXPECT(x.size() == 64 && x[0] == 1.f && x[2] == 0.25f && x[3] == 0.125f);
  threw = false;
  try {
    p.num_bins = 65;
    bp::configure(p);
  } catch (const xeno::error &) {
    threw = true;
  }
  EXPECT(threw);
  std::printf(fails ? "FAILED %d\n" : "OK\n", fails);
  return fails ? 1 : 0;
}
