// deep_agent -- the reference's known-answer evaluator (apps/bin_packing/deep_agent.cc): load the
// flat fp32 checkpoint `weights.20` (conv1d 4 -> 128 -> 64 -> 1, 8961 parameters), play the argmax
// policy, report the mean reward per episode (reference logs: 26.553 +- 0.028 per 10 000 episodes).
//   deep_agent <weights file> [num_envs] [episodes_per_env]
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <memory>
#include <vector>

#include <apps/bin_packing/bin_packing.h>

int main(int argc, char **argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: deep_agent <weights file> [num_envs] [episodes_per_env]\n");
    return 2;
  }
  const std::size_t num_envs = argc > 2 ? std::strtoul(argv[2], nullptr, 10) : 8192;
  const int episodes = argc > 3 ? std::atoi(argv[3]) : 4;

  xylo::model action_model;  // deep_agent.cc:11-17
  action_model.add_layer(std::make_unique<xylo::convolution1d_1_layer>(4, 128));
  action_model.add_layer(std::make_unique<xylo::relu_activation>());
  action_model.add_layer(std::make_unique<xylo::convolution1d_1_layer>(128, 64));
  action_model.add_layer(std::make_unique<xylo::relu_activation>());
  action_model.add_layer(std::make_unique<xylo::convolution1d_1_layer>(64, 1));
  action_model.bind_input_cols(bp::observation::length());

  std::ifstream f(argv[1], std::ios::binary);
  std::vector<float> weights(8961);
  f.read(reinterpret_cast<char *>(weights.data()), sizeof(float) * weights.size());
  if (f.gcount() != (std::streamsize)(sizeof(float) * weights.size())) {
    std::fprintf(stderr, "%s: expected %zu floats\n", argv[1], weights.size());
    return 1;
  }
  action_model.set_parameters(xylo::borrow_vector(weights));

  xylo::policy_gradient_deterministic_policy<bp::action, bp::observation> policy(action_model);
  bp::environment env(num_envs, 2021);
  xylo::replay_buffer<bp::action, bp::observation> rb;
  bp::agent agent(policy, env, rb);
  for (int i = 0; i < episodes; ++i)
    agent.play_one_episode();
  std::printf("mean_reward %.4f episodes %lld\n", xylo::total_rewards(rb) / double(rb.store().eval_episodes),
              rb.store().eval_episodes);
  return 0;
}
