// ppo_training -- the reference trainer main (apps/bin_packing/ppo_training.cc) on the device.
// Same phases: rollout (agent.play_steps) -> learner.step() -> replay_buffer.forget() -> periodic
// argmax evaluation. The 8 worker threads x 8 environments become one batched environment.
//   ppo_training [--num_bins=B] [--capacity=C] [--p_shape1=P] [num_envs] [iterations] [eval_every] [nets]
// nets = "ref" (default): the reference's own nets, line for line (ppo_training.cc:10-26: conv1d_1
// 4-128-64-1 softmax policy over the bins, critic 4B-64-32-1); "c2": the BASELINE configs[1] nets
// (policy 4B-64-64-B softmax, critic 4B-64-64-1). With the reference's 8 bins both run on the fused
// tcgen05 kernels. The flags (xeno::flagstore, reference xeno/configuration.h) turn the reference's
// compile-time problem definition (bin_packing.h:12, 24, 73-78) into launch parameters: e.g.
// `ppo_training --num_bins=32 4096 100 0 c2` is the "more bins" sweep without recompiling.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>

#include <xeno/configuration.h>
#include <xylo/nn.h>
#include <xylo/rl.h>

#include <apps/bin_packing/bin_packing.h>

int main(int argc, char **argv) {
  xeno::flagstore flags;
  flags.define_flag<int>("num_bins", 'b', 8);          // bin_packing.h:12
  flags.define_flag<int>("capacity", 'c', 8);          // bin_packing.h:24 (square bins)
  flags.define_flag<double>("p_shape1", 'p', 0.4);     // bin_packing.h:50
  const std::vector<std::string_view> pos = flags.parse_from_args(argc, argv);
  bp::problem problem;
  problem.num_bins = (std::size_t)flags.get_flag<int>("num_bins");
  problem.capacity = {flags.get_flag<int>("capacity"), flags.get_flag<int>("capacity")};
  problem.p_shape1 = (float)flags.get_flag<double>("p_shape1");
  bp::configure(problem);
  auto arg = [&](std::size_t i) { return std::string(pos[i]); };
  const std::size_t num_envs = pos.size() > 0 ? std::stoul(arg(0)) : 4096;
  const int iterations = pos.size() > 1 ? std::stoi(arg(1)) : 1000;
  const int eval_every = pos.size() > 2 ? std::stoi(arg(2)) : 100;
  constexpr int steps_per_worker = 4;  // ppo_training.cc:31
  // gradients are SUMS over rows (nn.h:94-98): the reference rates are tuned to 8 x 4 = 32 rows
  const float row_scale = 32.f / float(num_envs * steps_per_worker);

  const bool c2 = pos.size() > 3 && pos[3] == "c2";

  xylo::model action_model;
  xylo::model value_model;
  if (c2) {
    action_model.add_layer(std::make_unique<xylo::full_layer>(4 * bp::num_bins, 64));
    action_model.add_layer(std::make_unique<xylo::relu_activation>());
    action_model.add_layer(std::make_unique<xylo::full_layer>(64, 64));
    action_model.add_layer(std::make_unique<xylo::relu_activation>());
    action_model.add_layer(std::make_unique<xylo::full_layer>(64, bp::num_bins));
    action_model.add_layer(std::make_unique<xylo::softmax_layer>());
    value_model.add_layer(std::make_unique<xylo::full_layer>(4 * bp::num_bins, 64));
    value_model.add_layer(std::make_unique<xylo::relu_activation>());
    value_model.add_layer(std::make_unique<xylo::full_layer>(64, 64));
    value_model.add_layer(std::make_unique<xylo::relu_activation>());
    value_model.add_layer(std::make_unique<xylo::full_layer>(64, 1));
  } else {
    action_model.add_layer(std::make_unique<xylo::convolution1d_1_layer>(4, 128));
    action_model.add_layer(std::make_unique<xylo::relu_activation>());
    action_model.add_layer(std::make_unique<xylo::convolution1d_1_layer>(128, 64));
    action_model.add_layer(std::make_unique<xylo::relu_activation>());
    action_model.add_layer(std::make_unique<xylo::convolution1d_1_layer>(64, 1));
    action_model.add_layer(std::make_unique<xylo::softmax_layer>());
    value_model.add_layer(std::make_unique<xylo::full_layer>(4 * bp::num_bins, 64));
    value_model.add_layer(std::make_unique<xylo::relu_activation>());
    value_model.add_layer(std::make_unique<xylo::full_layer>(64, 32));
    value_model.add_layer(std::make_unique<xylo::relu_activation>());
    value_model.add_layer(std::make_unique<xylo::full_layer>(32, 1));
  }
  xylo::sgd_optimizer action_optimizer(action_model, 1e-4 * row_scale);
  xylo::sgd_optimizer value_optimizer(value_model, 1e-5 * row_scale);
  action_model.set_init_seed(1234);
  value_model.set_init_seed(1235);

  xylo::replay_buffer<bp::action, bp::observation> replay_buffer;
  bp::environment env(num_envs);
  xylo::policy_gradient_policy<bp::action, bp::observation> policy(action_model);
  bp::agent agent(policy, env, replay_buffer);
  bp::ppo_learner learner(replay_buffer, action_model, action_optimizer, value_model, value_optimizer, 0.99);

  auto t0 = std::chrono::steady_clock::now();
  for (int steps = 0; steps < iterations; ++steps) {
    agent.play_steps(steps_per_worker);
    learner.step();
    replay_buffer.forget();

    if (eval_every > 0 && steps % eval_every == 0) {
      xylo::policy_gradient_deterministic_policy<bp::action, bp::observation> greedy(action_model);
      bp::environment eval_env(1024, 99);
      xylo::replay_buffer<bp::action, bp::observation> rb;
      bp::agent eval_agent(greedy, eval_env, rb);
      eval_agent.play_one_episode();
      std::printf("round %d %.3f\n", steps, xylo::total_rewards(rb) / 1024.0);
    }
  }
  xylo::device::sync();
  double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  dfrl_trainer_stats s = replay_buffer.stats();
  xylo::vector p = action_model.parameters();
  double sum = 0;
  for (std::size_t i = 0; i < p.size(); ++i)
    sum += p[i];
  std::printf("env_steps %lld episodes %lld reward_sum %.0f param_sum %.9g env_steps_per_s %.3e\n", s.env_steps,
              s.episodes, s.reward_sum, sum, s.env_steps / secs);
  return 0;
}
