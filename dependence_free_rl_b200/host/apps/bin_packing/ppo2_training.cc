// ppo2_training -- the reference's KL-regulated PPO trainer (apps/bin_packing/ppo2_training.cc) on the device:
// kl_ppo_learner (policy_gradient.h:310-335: k = 4 steps of A (p - onehot) + beta (p - p_old) through the softmax
// Jacobian, beta adapted around d_targ after every step -- on the device, no host round trip), SGD with weight
// decay 1e-5 on the policy (ppo2_training.cc:20), 8 steps per environment and iteration (:34).
//   ppo2_training [num_envs] [iterations] [eval_every]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>

#include <apps/bin_packing/bin_packing.h>

int main(int argc, char **argv) {
  const std::size_t num_envs = argc > 1 ? std::strtoul(argv[1], nullptr, 10) : 4096;
  const int iterations = argc > 2 ? std::atoi(argv[2]) : 200;
  const int eval_every = argc > 3 ? std::atoi(argv[3]) : 100;
  constexpr int steps_per_worker = 8;
  const float row_scale = 128.f / float(num_envs * steps_per_worker);  // reference: 16 workers x 8 steps (gradients are sums)

  xylo::model action_model;  // ppo2_training.cc:12-19
  action_model.add_layer(std::make_unique<xylo::convolution1d_1_layer>(4, 128));
  action_model.add_layer(std::make_unique<xylo::relu_activation>());
  action_model.add_layer(std::make_unique<xylo::convolution1d_1_layer>(128, 64));
  action_model.add_layer(std::make_unique<xylo::relu_activation>());
  action_model.add_layer(std::make_unique<xylo::convolution1d_1_layer>(64, 1));
  action_model.add_layer(std::make_unique<xylo::softmax_layer>());
  xylo::sgd_optimizer action_optimizer(action_model, 1e-4f * row_scale, 1e-5f);

  xylo::model value_model;  // ppo2_training.cc:22-29
  value_model.add_layer(std::make_unique<xylo::full_layer>(4 * bp::num_bins, 64));
  value_model.add_layer(std::make_unique<xylo::relu_activation>());
  value_model.add_layer(std::make_unique<xylo::full_layer>(64, 32));
  value_model.add_layer(std::make_unique<xylo::relu_activation>());
  value_model.add_layer(std::make_unique<xylo::full_layer>(32, 1));
  xylo::sgd_optimizer value_optimizer(value_model, 1e-5f * row_scale);
  action_model.set_init_seed(1234);
  value_model.set_init_seed(1235);

  xylo::replay_buffer<bp::action, bp::observation> replay_buffer;
  bp::environment env(num_envs);
  xylo::policy_gradient_policy<bp::action, bp::observation> policy(action_model);
  bp::agent agent(policy, env, replay_buffer);
  bp::kl_ppo_learner learner(replay_buffer, action_model, action_optimizer, value_model, value_optimizer, 0.99);

  const auto t0 = std::chrono::steady_clock::now();
  for (int steps = 0; steps < iterations; ++steps) {
    agent.play_steps(steps_per_worker);
    learner.step();
    replay_buffer.forget();
    if (eval_every > 0 && steps % eval_every == 0) {  // ppo2_training.cc:75-92
      xylo::policy_gradient_deterministic_policy<bp::action, bp::observation> greedy(action_model);
      bp::environment eval_env(1024, 99);
      xylo::replay_buffer<bp::action, bp::observation> rb;
      bp::agent eval_agent(greedy, eval_env, rb);
      eval_agent.play_one_episode();
      std::printf("round %d %.3f\n", steps, xylo::total_rewards(rb) / 1024.0);
    }
  }
  xylo::device::sync();
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  const dfrl_trainer_stats s = replay_buffer.stats();
  std::printf("env_steps %lld episodes %lld reward_sum %.0f kl_beta %.3g env_steps_per_s %.3e\n", s.env_steps, s.episodes,
              s.reward_sum, (double)s.kl_beta, s.env_steps / secs);
  return 0;
}
