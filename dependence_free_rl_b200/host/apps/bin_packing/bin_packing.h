// apps/bin_packing/bin_packing.h -- the bin-packing problem of the reference
// (apps/bin_packing/bin_packing.h:12-149) on the batched device environment.
#ifndef BIN_PACKING
#define BIN_PACKING

#include <sstream>
#include <string>
#include <utility>
#include <vector>

#include <xylo/nn.h>
#include <xylo/policy_gradient.h>

namespace bp {

// The problem definition. The reference fixes it at compile time (bin_packing.h:12 `constexpr
// std::size_t num_bins = 8`, :24 capacity (8, 8), :73-78 items (4, 2) with probability 0.4 else
// (1, 2)); here it is a set of LAUNCH parameters of the device environment (dfrl_env_config): set
// them before constructing models / environments (the trainer mains do it from command-line flags,
// xeno/configuration.h) and sweep the number of bins without recompiling.
struct problem {
  std::size_t num_bins = 8;
  std::pair<int, int> capacity{8, 8};
  std::pair<int, int> shape1{4, 2}, shape2{1, 2};
  float p_shape1 = 0.4f;
};
inline problem &config() {
  static problem p;
  return p;
}
inline void configure(const problem &p) {
  if (p.num_bins < 1 || p.num_bins > 64)
    throw xeno::error("num_bins must lie in [1, 64]");
  config() = p;
  xylo::dynamic_cardinality() = p.num_bins;
}
// `bp::num_bins` reads like the reference's constant (`4 * bp::num_bins`, `bins.size() == bp::num_bins`).
inline constexpr struct num_bins_t {
  operator std::size_t() const { return config().num_bins; }
} num_bins;

using action = xylo::discrete_action<0>;  // 0: run-time cardinality (= num_bins, set by configure())

struct observation {
  static std::size_t length() { return 4 * config().num_bins; }

  observation(const std::pair<int, int> &bin_shape) : bins(config().num_bins, bin_shape), item{0, 0} {}

  std::string to_string() const {
    std::ostringstream oss;
    oss << "item: (" << item.first << ", " << item.second << "); bins:";
    for (const auto &b : bins)
      oss << " (" << b.first << ", " << b.second << ")";
    return oss.str();
  }

  // [bin.w / cap.w, bin.h / cap.h, item.w / cap.w, item.h / cap.h] per bin (bin_packing.h:31-40)
  void to_vector(xylo::vector_view o) const {
    const std::pair<int, int> capacity = config().capacity;
    for (std::size_t i = 0; i < bins.size(); ++i) {
      o[4 * i + 0] = float(bins[i].first) / capacity.first;
      o[4 * i + 1] = float(bins[i].second) / capacity.second;
      o[4 * i + 2] = float(item.first) / capacity.first;
      o[4 * i + 3] = float(item.second) / capacity.second;
    }
  }

  std::vector<std::pair<int, int>> bins;
  std::pair<int, int> item;
};

// N independent environments in HBM; `id` selects one. environment() is the reference's single
// environment (bin_packing.h:50-52).
class environment : public xylo::environment<action, observation> {
public:
  explicit environment(std::size_t n_envs = 1, uint64_t seed = 1234, int64_t env_offset = 0) : n_(n_envs), cfg_(config()) {
    dfrl_env_config c;
    dfrl_env_config_default(&c);  // 8 bins of (8, 8); items (4, 2) w.p. 0.4 else (1, 2)
    c.n_envs = (int)n_envs;
    c.n_bins = (int)cfg_.num_bins;
    c.cap_w = cfg_.capacity.first;
    c.cap_h = cfg_.capacity.second;
    c.item_w[0] = cfg_.shape1.first, c.item_h[0] = cfg_.shape1.second;
    c.item_w[1] = cfg_.shape2.first, c.item_h[1] = cfg_.shape2.second;
    c.p_shape1 = cfg_.p_shape1;
    c.seed = seed;
    c.env_offset = env_offset;
    xylo::check(dfrl_env_create(xylo::device::get(), &c, &env_));
  }
  environment(const environment &) = delete;
  environment(environment &&o) noexcept : n_(o.n_), cfg_(o.cfg_), env_(o.env_) { o.env_ = nullptr; }
  ~environment() override {
    if (env_)
      dfrl_env_destroy(env_);
  }

  void apply(const action &action, std::size_t id) override {
    xylo::check(dfrl_env_apply_one(env_, (int)id, (int)action.choice));
  }
  observation view(std::size_t id) const override {
    const std::size_t B = cfg_.num_bins;
    std::vector<int8_t> s(2 * B + 2);
    xylo::check(dfrl_env_view_one(env_, (int)id, s.data()));
    observation o(cfg_.capacity);
    for (std::size_t b = 0; b < B; ++b)
      o.bins[b] = {s[2 * b], s[2 * b + 1]};
    o.item = {s[2 * B], s[2 * B + 1]};
    return o;
  }
  void reset(std::size_t id) override { xylo::check(dfrl_env_reset_one(env_, (int)id)); }

  dfrl_env *device_env() override { return env_; }
  std::size_t size() const override { return n_; }

private:
  std::size_t n_;
  problem cfg_;  // the problem definition this environment was created with
  dfrl_env *env_ = nullptr;
};

class agent : public xylo::agent<action, observation> {
public:
  agent(const xylo::policy<action, observation> &p, environment &env, xylo::replay_buffer<action, observation> &rb)
      : xylo::agent<action, observation>(p, env, rb) {}

private:
  bool game_over(const observation &ob) override {
    for (const auto &bin : ob.bins)
      if (bin.first < 0 || bin.second < 0)
        return true;
    return false;
  }
  float get_reward(const observation &, const observation &ob) override { return game_over(ob) ? 0 : 1; }
};

// The reference's rule-based agents (firstfit_agent.cc:10-28, bestfit_agent.cc:10-30, minwaste_agent.cc:10-39,
// random_agent.cc + rl.h:305-315): there each main defines a policy class with a host react(); here the four
// rules are built into the environment kernel (dfrl_heuristic_react / dfrl_heuristic_play) and these classes
// name them. react() is the per-state slow path: the state goes through the SAME device rule on a private
// one-slot environment.
class rule_policy : public xylo::policy<action, observation> {
public:
  explicit rule_policy(int kind) : kind_(kind) {}
  rule_policy(const rule_policy &) = delete;
  ~rule_policy() override {
    if (one_)
      dfrl_env_destroy(one_);
  }
  int device_rule() const override { return kind_; }
  action react(const observation &state) const override {
    const std::size_t B = config().num_bins;
    if (state.bins.size() != B)
      throw xeno::error("observation with the wrong number of bins");
    if (!one_) {
      dfrl_env_config c;
      dfrl_env_config_default(&c);
      c.n_envs = 1;
      c.n_bins = (int)B;
      c.cap_w = config().capacity.first, c.cap_h = config().capacity.second;
      xylo::check(dfrl_env_create(xylo::device::get(), &c, &one_));
    }
    std::vector<int8_t> planes(2 * B + 2);  // host layout of dfrl_env_set_state: [2B + 2][N], N = 1
    for (std::size_t b = 0; b < B; ++b) {
      planes[2 * b] = (int8_t)state.bins[b].first;
      planes[2 * b + 1] = (int8_t)state.bins[b].second;
    }
    planes[2 * B] = (int8_t)state.item.first;
    planes[2 * B + 1] = (int8_t)state.item.second;
    xylo::check(dfrl_env_set_state(one_, planes.data()));
    xylo::device_buffer act(1);
    xylo::check(dfrl_heuristic_react(one_, kind_, reinterpret_cast<uint8_t *>(act.get())));
    uint8_t c = 0;
    xylo::check(dfrl_memcpy_d2h(xylo::device::get(), &c, act.get(), 1));
    action a;
    a.choice = c;
    return a;
  }

private:
  int kind_;
  mutable dfrl_env *one_ = nullptr;
};
struct firstfit_policy : rule_policy { firstfit_policy() : rule_policy(DFRL_HEUR_FIRSTFIT) {} };
struct bestfit_policy : rule_policy { bestfit_policy() : rule_policy(DFRL_HEUR_BESTFIT) {} };
struct minwaste_policy : rule_policy { minwaste_policy() : rule_policy(DFRL_HEUR_MINWASTE) {} };
struct random_policy : rule_policy { random_policy() : rule_policy(DFRL_HEUR_RANDOM) {} };

class pg_learner : public xylo::policy_gradient_learner<action, observation> {
public:
  pg_learner(xylo::replay_buffer<action, observation> &rb, xylo::model &action_model,
             xylo::optimizer &action_optimizer, float gamma = 0.99)
      : xylo::policy_gradient_learner<action, observation>(rb, action_model, action_optimizer, gamma) {}
};

class ac_learner : public xylo::actor_critic_learner<action, observation> {
public:
  ac_learner(xylo::replay_buffer<action, observation> &rb, xylo::model &action_model,
             xylo::optimizer &action_optimizer, xylo::model &value_model, xylo::optimizer &value_optimizer,
             float gamma = 0.99)
      : xylo::actor_critic_learner<action, observation>(rb, action_model, action_optimizer, value_model,
                                                        value_optimizer, gamma) {}
};

class ppo_learner : public xylo::ppo_learner<action, observation> {
public:
  ppo_learner(xylo::replay_buffer<action, observation> &rb, xylo::model &action_model,
              xylo::optimizer &action_optimizer, xylo::model &value_model, xylo::optimizer &value_optimizer,
              float gamma = 0.99)
      : xylo::ppo_learner<action, observation>(rb, action_model, action_optimizer, value_model, value_optimizer,
                                               gamma) {}
};

class kl_ppo_learner : public xylo::kl_ppo_learner<action, observation> {
public:
  kl_ppo_learner(xylo::replay_buffer<action, observation> &rb, xylo::model &action_model,
                 xylo::optimizer &action_optimizer, xylo::model &value_model, xylo::optimizer &value_optimizer,
                 float gamma = 0.99)
      : xylo::kl_ppo_learner<action, observation>(rb, action_model, action_optimizer, value_model, value_optimizer,
                                                  gamma) {}
};

} // namespace bp

// How replay_buffer::sample_td() (xylo/rl.h) rebuilds observations from the device's rollout record: one column
// of 2B + 2 int8 planes [bin0.w, bin0.h, ..., item.w, item.h]; a finished episode ends in the overflowed state
// (the chosen bin minus the item, item kept: reference bin_packing.h:54-61), which the device derives instead
// of storing.
namespace xylo {
template <> struct record_codec<bp::observation> {
  static bp::observation decode(const int8_t *s) {
    const std::size_t B = bp::config().num_bins;
    bp::observation o(bp::config().capacity);
    for (std::size_t b = 0; b < B; ++b)
      o.bins[b] = {s[2 * b], s[2 * b + 1]};
    o.item = {s[2 * B], s[2 * B + 1]};
    return o;
  }
  static bp::observation terminal(const bp::observation &start, std::size_t choice) {
    bp::observation o = start;
    o.bins[choice].first -= o.item.first;
    o.bins[choice].second -= o.item.second;
    return o;
  }
};
} // namespace xylo

#endif
