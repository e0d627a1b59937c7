// xylo/policy_gradient.h -- the learners and NN policies of the reference (xylo/policy_gradient.h:
// 89-373) with their constructor signatures; learn() runs the fused device iteration
// (dfrl_trainer_learn): returns / GAE, loss gradients, backward passes and optimizer updates.
#ifndef XYLO_POLICY_GRADIENT_
#define XYLO_POLICY_GRADIENT_

#include <xylo/rl.h>

namespace xylo {

// REINFORCE (policy_gradient.h:89-148): reversed-order discounted returns, trajectory-mean baseline.
template <typename A, typename S> class policy_gradient_learner : public learner<A, S> {
public:
  policy_gradient_learner(replay_buffer<A, S> &rb, model &action_model, optimizer &action_optimizer,
                          float gamma = 0.99)
      : learner<A, S>(rb, action_model, action_optimizer, gamma) {
    attach(rb, DFRL_ALGO_REINFORCE, action_model, action_optimizer, nullptr, nullptr, gamma);
  }
  void learn() override { this->learn_on_device(); }

protected:
  static void attach(replay_buffer<A, S> &rb, int algo, model &am, optimizer &ao, model *vm, optimizer *vo,
                     float gamma) {
    rollout_store &s = rb.store();
    s.algo = algo;
    s.policy_model = &am;
    s.policy_opt = &ao;
    s.value_model = vm;
    s.value_opt = vo;
    s.gamma = gamma;
    s.drop();
  }
};

// Online actor-critic (policy_gradient.h:150-287): TD(0) critic step, GAE(lambda = 0.95) from the
// updated critic, one policy step with A (p - onehot).
template <typename A, typename S> class actor_critic_learner : public policy_gradient_learner<A, S> {
public:
  actor_critic_learner(replay_buffer<A, S> &rb, model &action_model, optimizer &action_optimizer,
                       model &value_model, optimizer &value_optimizer, float gamma = 0.99)
      : policy_gradient_learner<A, S>(rb, action_model, action_optimizer, gamma) {
    this->attach(rb, DFRL_ALGO_ACTOR_CRITIC, action_model, action_optimizer, &value_model, &value_optimizer, gamma);
  }

protected:
  void retag(int algo) { this->replay_buffer_.store().algo = algo; }
};

// PPO-clip (policy_gradient.h:289-308): k = 4 full-batch steps of the clipped surrogate.
template <typename A, typename S> class ppo_learner : public actor_critic_learner<A, S> {
public:
  ppo_learner(replay_buffer<A, S> &rb, model &action_model, optimizer &action_optimizer, model &value_model,
              optimizer &value_optimizer, float gamma = 0.99)
      : actor_critic_learner<A, S>(rb, action_model, action_optimizer, value_model, value_optimizer, gamma) {
    this->retag(DFRL_ALGO_PPO);
  }
};

// PPO with adaptive KL penalty (policy_gradient.h:310-335).
template <typename A, typename S> class kl_ppo_learner : public actor_critic_learner<A, S> {
public:
  kl_ppo_learner(replay_buffer<A, S> &rb, model &action_model, optimizer &action_optimizer, model &value_model,
                 optimizer &value_optimizer, float gamma = 0.99)
      : actor_critic_learner<A, S>(rb, action_model, action_optimizer, value_model, value_optimizer, gamma) {
    this->retag(DFRL_ALGO_KL_PPO);
  }
};

// policy_gradient_policy (policy_gradient.h:337-353): sample from the net's output distribution.
template <typename A, typename S> class policy_gradient_policy : public policy<A, S> {
public:
  explicit policy_gradient_policy(model &m) : model_(m) {}
  A react(const S &state) const override {
    vector in = to_vector(state);
    matrix out = model_.eval(matrix_view(in.data(), {1, in.size()}));
    A a;
    a.from_vector(out[0]);
    return a;
  }
  model *backing_model() const override { return &model_; }

protected:
  model &model_;
};

// policy_gradient_deterministic_policy (policy_gradient.h:355-372): argmax, first maximum wins.
template <typename A, typename S> class policy_gradient_deterministic_policy : public policy_gradient_policy<A, S> {
public:
  using policy_gradient_policy<A, S>::policy_gradient_policy;
  A react(const S &state) const override {
    vector in = to_vector(state);
    matrix out = this->model_.eval(matrix_view(in.data(), {1, in.size()}));
    A a;
    a.from_vector_deterministic(out[0]);
    return a;
  }
  bool deterministic() const override { return true; }
};

} // namespace xylo

#endif // XYLO_POLICY_GRADIENT_
