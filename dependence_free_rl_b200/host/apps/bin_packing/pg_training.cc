// pg_training -- REINFORCE (apps/bin_packing/pg_training.cc): every environment plays whole
// episodes, one policy step on the reversed-order returns minus the trajectory-mean baseline.
//   pg_training [num_envs] [iterations]
#include <cstdio>
#include <cstdlib>
#include <memory>

#include <apps/bin_packing/bin_packing.h>

int main(int argc, char **argv) {
  const std::size_t num_envs = argc > 1 ? std::strtoul(argv[1], nullptr, 10) : 1024;
  const int iterations = argc > 2 ? std::atoi(argv[2]) : 100;
  const float row_scale = 16.f / float(num_envs);  // reference: 4 workers x 4 episodes per iteration

  xylo::model action_model;  // pg_training.cc:11-18
  action_model.add_layer(std::make_unique<xylo::full_layer>(4 * bp::num_bins, 256));
  action_model.add_layer(std::make_unique<xylo::relu_activation>());
  action_model.add_layer(std::make_unique<xylo::full_layer>(256, 128));
  action_model.add_layer(std::make_unique<xylo::relu_activation>());
  action_model.add_layer(std::make_unique<xylo::full_layer>(128, bp::num_bins));
  action_model.add_layer(std::make_unique<xylo::softmax_cross_entropy_layer>());
  xylo::sgd_optimizer action_optimizer(action_model, 1e-4 * row_scale);

  xylo::replay_buffer<bp::action, bp::observation> replay_buffer;
  bp::environment env(num_envs);
  xylo::policy_gradient_policy<bp::action, bp::observation> policy(action_model);
  bp::agent agent(policy, env, replay_buffer);
  bp::pg_learner learner(replay_buffer, action_model, action_optimizer, 0.99);

  for (int steps = 0; steps < iterations; ++steps) {
    agent.play_one_episode();
    learner.step();
    replay_buffer.forget();
  }
  dfrl_trainer_stats s = replay_buffer.stats();
  std::printf("env_steps %lld episodes %lld reward_sum %.0f mean_episode_reward %.3f\n", s.env_steps, s.episodes,
              s.reward_sum, s.episodes ? s.reward_sum / s.episodes : 0.0);
  return 0;
}
