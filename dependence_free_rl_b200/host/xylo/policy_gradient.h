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
//
// optimize_action (policy_gradient.h:187-194) is the reference's specialisation point: ppo_learner
// (297-307) and kl_ppo_learner (318-330) override it, and so may a user's learner. Here the built-in
// learners share ONE default (the device policy phase of the attached algorithm), and learn() asks
// whether the final overrider is that default:
//   * it is  -> the whole learn() runs on the device (dfrl_trainer_learn: CUDA-graphed kernels);
//   * a user override -> critic step + advantages on the device (dfrl_trainer_learn_phases), then
//     the override is called with what the reference passes it -- the start-state observations
//     [rows][O], the recorded actions (choice + old distribution) and the advantages -- on the host.
//     (The reference also passes one zero-advantage end-state row per trajectory, quirk 8; those rows
//     contribute nothing to any gradient and are not materialised.)
// The test needs g++'s bound-member-function conversion (-Wno-pmf-conversions); other compilers
// opt in with use_optimize_action_hook(true) in the derived constructor.
template <typename A, typename S> class actor_critic_learner : public policy_gradient_learner<A, S> {
public:
  actor_critic_learner(replay_buffer<A, S> &rb, model &action_model, optimizer &action_optimizer,
                       model &value_model, optimizer &value_optimizer, float gamma = 0.99)
      : policy_gradient_learner<A, S>(rb, action_model, action_optimizer, gamma) {
    this->attach(rb, DFRL_ALGO_ACTOR_CRITIC, action_model, action_optimizer, &value_model, &value_optimizer, gamma);
  }

  void learn() override {
    if (!optimize_action_overridden()) {
      this->learn_on_device();
      return;
    }
    rollout_store &s = this->replay_buffer_.store();
    if (!s.trainer || !s.rolled)
      throw xeno::error("learn() before any rollout");
    // update_value_model (196-218) + calculate_advantage (220-281) with the updated critic
    s.sync_rates();
    s.learned = true;
    check(dfrl_trainer_learn_phases(s.trainer, DFRL_PHASE_VALUE | DFRL_PHASE_ADVANTAGE));
    std::size_t bytes = 0;
    check(dfrl_trainer_field_size(s.trainer, DFRL_F_ADVANTAGE, &bytes));
    const std::size_t rows = bytes / sizeof(float), B = A::cardinality();
    matrix states({rows, s.obs_cols});
    vector advantages({rows});
    std::vector<uint8_t> choice(rows);
    std::vector<float> probs(rows * B);
    check(dfrl_trainer_read(s.trainer, DFRL_F_OBS_START, states.data(), sizeof(float) * states.size()));
    check(dfrl_trainer_read(s.trainer, DFRL_F_ADVANTAGE, advantages.data(), bytes));
    check(dfrl_trainer_read(s.trainer, DFRL_F_REC_ACTION, choice.data(), rows));
    check(dfrl_trainer_read(s.trainer, DFRL_F_REC_PROBS, probs.data(), sizeof(float) * probs.size()));
    std::vector<A> actions(rows);
    for (std::size_t i = 0; i < rows; ++i) {
      actions[i].choice = choice[i];
      vector d({B});
      for (std::size_t q = 0; q < B; ++q)
        d[q] = probs[i * B + q];
      actions[i].distrib = std::move(d);
    }
    optimize_action(states.view(), actions, advantages.view());
  }

protected:
  // Default of every built-in learner: the device policy phase -- one A (p - onehot) step, k clipped
  // steps or k KL-regulated steps, by the learner's algorithm. Its arguments are not touched.
  virtual void optimize_action(matrix_view /*states*/, const std::vector<A> & /*actions*/, vector_view /*advantages*/) {
    rollout_store &s = this->replay_buffer_.store();
    check(dfrl_trainer_learn_phases(s.trainer, DFRL_PHASE_POLICY));
  }
  void use_optimize_action_hook(bool on) { hook_forced_ = on; }
  void retag(int algo) { this->replay_buffer_.store().algo = algo; }

private:
  bool optimize_action_overridden() {
    if (hook_forced_)
      return true;
#if defined(__GNUC__) && !defined(__clang__)
#pragma GCC diagnostic push
#pragma GCC diagnostic ignored "-Wpmf-conversions"
    typedef void (*raw_t)(actor_critic_learner *, matrix_view, const std::vector<A> &, vector_view);
    const raw_t final_overrider = (raw_t)(this->*(&actor_critic_learner::optimize_action));
    const raw_t builtin = (raw_t)(&actor_critic_learner::optimize_action);
#pragma GCC diagnostic pop
    return final_overrider != builtin;
#else
    return false;
#endif
  }
  bool hook_forced_ = false;
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
