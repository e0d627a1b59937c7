// ppo_training -- the reference trainer main (apps/bin_packing/ppo_training.cc) on the device.
// Same phases: rollout (agent.play_steps) -> learner.step() -> replay_buffer.forget() -> periodic
// argmax evaluation. The 8 worker threads x 8 environments become one batched environment.
//   ppo_training [num_envs] [iterations] [eval_every] [nets]
// nets = "ref" (default): the reference's own nets, line for line (ppo_training.cc:10-26: conv1d_1
// 4-128-64-1 softmax policy over the 8 bins, critic 32-64-32-1); "c2": the BASELINE configs[1] nets
// (policy 32-64-64-8 softmax, critic 32-64-64-1). Both run on the fused tcgen05 kernels.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>

#include <xylo/nn.h>
#include <xylo/rl.h>

#include <apps/bin_packing/bin_packing.h>

int main(int argc, char **argv) {
  const std::size_t num_envs = argc > 1 ? std::strtoul(argv[1], nullptr, 10) : 4096;
  const int iterations = argc > 2 ? std::atoi(argv[2]) : 1000;
  const int eval_every = argc > 3 ? std::atoi(argv[3]) : 100;
  constexpr int steps_per_worker = 4;  // ppo_training.cc:31
  // gradients are SUMS over rows (nn.h:94-98): the reference rates are tuned to 8 x 4 = 32 rows
  const float row_scale = 32.f / float(num_envs * steps_per_worker);

  const bool c2 = argc > 4 && std::strcmp(argv[4], "c2") == 0;

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
