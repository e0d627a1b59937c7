// ac_training -- online actor-critic (apps/bin_packing/ac_training.cc): play_steps(8), one critic
// step, GAE from the updated critic, one policy step with A (p - onehot).
//   ac_training [num_envs] [iterations]
#include <cstdio>
#include <cstdlib>
#include <memory>

#include <apps/bin_packing/bin_packing.h>

int main(int argc, char **argv) {
  const std::size_t num_envs = argc > 1 ? std::strtoul(argv[1], nullptr, 10) : 65536;
  const int iterations = argc > 2 ? std::atoi(argv[2]) : 100;
  constexpr int steps_per_worker = 8;  // ac_training.cc:30
  const float row_scale = 128.f / float(num_envs * steps_per_worker);  // reference: 16 workers x 8 steps

  xylo::model action_model;  // per-bin shared weights (conv1d over the 8 bins), ac_training.cc:9-16
  action_model.add_layer(std::make_unique<xylo::convolution1d_1_layer>(4, 64));
  action_model.add_layer(std::make_unique<xylo::relu_activation>());
  action_model.add_layer(std::make_unique<xylo::convolution1d_1_layer>(64, 32));
  action_model.add_layer(std::make_unique<xylo::relu_activation>());
  action_model.add_layer(std::make_unique<xylo::convolution1d_1_layer>(32, 1));
  action_model.add_layer(std::make_unique<xylo::softmax_cross_entropy_layer>());
  xylo::sgd_optimizer action_optimizer(action_model, 1e-5 * row_scale);

  xylo::model value_model;
  value_model.add_layer(std::make_unique<xylo::full_layer>(4 * bp::num_bins, 64));
  value_model.add_layer(std::make_unique<xylo::relu_activation>());
  value_model.add_layer(std::make_unique<xylo::full_layer>(64, 32));
  value_model.add_layer(std::make_unique<xylo::relu_activation>());
  value_model.add_layer(std::make_unique<xylo::full_layer>(32, 1));
  xylo::sgd_optimizer value_optimizer(value_model, 1e-4 * row_scale);

  xylo::replay_buffer<bp::action, bp::observation> replay_buffer;
  bp::environment env(num_envs);
  xylo::policy_gradient_policy<bp::action, bp::observation> policy(action_model);
  bp::agent agent(policy, env, replay_buffer);
  bp::ac_learner learner(replay_buffer, action_model, action_optimizer, value_model, value_optimizer, 0.99);

  for (int steps = 0; steps < iterations; ++steps) {
    agent.play_steps(steps_per_worker);
    learner.step();
    replay_buffer.forget();
  }
  dfrl_trainer_stats s = replay_buffer.stats();
  std::printf("env_steps %lld episodes %lld reward_sum %.0f\n", s.env_steps, s.episodes, s.reward_sum);
  return 0;
}
