// heuristic_agent -- the reference's rule-based agents (apps/bin_packing/firstfit_agent.cc, bestfit_agent.cc,
// minwaste_agent.cc, random_agent.cc) in one main: every environment slot plays whole episodes under the rule,
// the mean reward per episode is printed per round (reference logs: min-waste 26.553 +- 0.009 per 100 000
// episodes, first-fit ~ 25.83, best-fit ~ 25.81, random ~ 11.6).
//   heuristic_agent <firstfit|bestfit|minwaste|random> [num_envs] [episodes_per_env] [rounds]
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>

#include <apps/bin_packing/bin_packing.h>

int main(int argc, char **argv) {
  const std::string rule = argc > 1 ? argv[1] : "minwaste";
  const std::size_t num_envs = argc > 2 ? std::strtoul(argv[2], nullptr, 10) : 16384;
  const int episodes = argc > 3 ? std::atoi(argv[3]) : 4;
  const int rounds = argc > 4 ? std::atoi(argv[4]) : 3;
  std::unique_ptr<bp::rule_policy> policy;
  if (rule == "firstfit")
    policy = std::make_unique<bp::firstfit_policy>();
  else if (rule == "bestfit")
    policy = std::make_unique<bp::bestfit_policy>();
  else if (rule == "minwaste")
    policy = std::make_unique<bp::minwaste_policy>();
  else if (rule == "random")
    policy = std::make_unique<bp::random_policy>();
  else {
    std::fprintf(stderr, "usage: heuristic_agent <firstfit|bestfit|minwaste|random> [num_envs] [episodes_per_env] [rounds]\n");
    return 2;
  }
  for (int round = 0; round < rounds; ++round) {  // minwaste_agent.cc:43-58
    bp::environment env(num_envs, 2021 + round);
    xylo::replay_buffer<bp::action, bp::observation> rb;
    bp::agent agent(*policy, env, rb);
    for (int i = 0; i < episodes; ++i)
      agent.play_one_episode();
    std::printf("round %d mean_reward %.4f episodes %lld\n", round, xylo::total_rewards(rb) / double(rb.store().eval_episodes),
                rb.store().eval_episodes);
  }
  // the per-state slow path goes through the same device rule: an empty board with a (4, 2) item
  bp::observation fresh(bp::config().capacity);
  fresh.item = bp::config().shape1;
  std::printf("react(empty board) -> bin %zu\n", policy->react(fresh).choice);
  return 0;
}
