// xylo/rl.h -- environment / agent / policy / learner / replay_buffer with the reference's
// signatures (xylo/rl.h:22-392), batched on the device.
//
// What changes underneath: the reference runs one environment object per agent and one pthread per
// agent (ppo_training.cc:28-61). Here ONE environment object holds N independent slots in HBM and
// `id` (already a parameter of environment::apply / view / reset, rl.h:163-170) selects a slot;
// ONE agent drives every slot, so agent::step() is one transition of every environment and
// play_steps(n) is the whole rollout phase of a trainer iteration in one launch. The replay buffer
// is the struct-of-arrays rollout record on the device (dfrl_trainer); sample_td() (rl.h:222-234) reads the
// last rollout's records back into the reference's transition / trajectory / td types (an inspection path:
// the learners never touch it), forget() is implicit (the next rollout overwrites the records and open
// trajectories continue from the live state, rl.h:274-291).
#ifndef XYLO_RL_
#define XYLO_RL_

#include <list>
#include <optional>
#include <random>
#include <vector>

#include <xylo/nn.h>
#include <xylo/tensor.h>

namespace xylo {

// The reference's global generator (tensor.h:17, tensor.cc:71-75): only the per-id slow path draws
// from it (one uniform per sampled action); batched rollouts use Philox keyed by (seed, env, step).
inline std::minstd_rand0 &default_generator() {
  static std::minstd_rand0 g(std::random_device{}());
  return g;
}

template <typename T> vector to_vector(const T &t) {
  vector result({t.length()});
  t.to_vector(result);
  return result;
}

// discrete_action<0>: the cardinality is a run-time value (apps/bin_packing: the number of bins as a launch parameter).
inline std::size_t &dynamic_cardinality() {
  static std::size_t n = 8;
  return n;
}

template <std::size_t range> struct discrete_action {
  static std::size_t cardinality() { return range ? range : dynamic_cardinality(); }
  std::size_t choice = 0;
  std::optional<vector> distrib;

  // std::discrete_distribution semantics (tensor.cc:467-470), evaluated by dfrl_sample.
  void from_vector(vector_view a) {
    const std::size_t range_ = cardinality();
    if (a.size() != range_)
      throw xeno::error("action vector has the wrong size");
    double u = std::generate_canonical<double, 53>(default_generator());
    device_buffer p(range_), ud(2), act(1);
    p.upload(a.data(), range_);
    check(dfrl_memcpy_h2d(device::get(), ud.get(), &u, sizeof(double)));
    check(dfrl_sample(device::get(), p.get(), 1, (int)range_, reinterpret_cast<const double *>(ud.get()),
                      reinterpret_cast<uint8_t *>(act.get()), nullptr));
    uint8_t c = 0;
    check(dfrl_memcpy_d2h(device::get(), &c, act.get(), 1));
    choice = c;
    distrib = vector(a);
  }
  void from_vector_deterministic(vector_view a) { choice = argmax(a); }

  // Loss-gradient rules of one row (rl.h:45-74), evaluated by dfrl_loss_grad.
  void softmax_gradient_log(vector_view input, vector_view output, float advantage) const {
    row_rule(DFRL_LOSS_SOFTMAX_LOG, input, output, advantage);
  }
  void clipped_gradient(vector_view input, vector_view output, float advantage) const {
    row_rule(DFRL_LOSS_CLIPPED, input, output, advantage);
  }

private:
  void row_rule(int kind, vector_view input, vector_view output, float advantage) const {
    const std::size_t range_ = cardinality();
    if (input.size() != range_ || output.size() != range_)
      throw xeno::error("action vector has the wrong size");
    device_buffer p(range_), o(range_), adv(1), act(1), po(1);
    p.upload(input.data(), range_);
    adv.upload(&advantage, 1);
    uint8_t c = (uint8_t)choice;
    check(dfrl_memcpy_h2d(device::get(), act.get(), &c, 1));
    float pold = distrib ? (*distrib)[choice] : 1.f;
    po.upload(&pold, 1);
    check(dfrl_loss_grad(device::get(), kind, p.get(), reinterpret_cast<const uint8_t *>(act.get()), adv.get(),
                         po.get(), 0.f, 1, (int)range_, o.get()));
    o.download(output.data(), range_);
  }
};

template <typename A, typename S> class environment {
public:
  virtual ~environment() = default;

  virtual void apply(const A &action, std::size_t id) = 0;
  virtual S view(std::size_t id) const = 0;
  virtual void reset(std::size_t id) = 0;

  // Batched extension: the device object holding every slot, and how many there are.
  virtual dfrl_env *device_env() { return nullptr; }
  virtual std::size_t size() const { return 1; }
};

template <typename A, typename S> class policy {
public:
  virtual ~policy() = default;

  virtual A react(const S &state) const = 0;

  // NN-backed policies run inside the rollout kernel: the net, and sample vs argmax.
  virtual model *backing_model() const { return nullptr; }
  virtual bool deterministic() const { return false; }
  // Rule-based policies (the reference's first-fit / best-fit / min-waste / random agents: user classes with a
  // host react()) run inside the environment kernel as one of its built-in rules: DFRL_HEUR_*, -1 = none.
  virtual int device_rule() const { return -1; }
};

// What the agent (rollout side) and the learner (update side) share through the replay buffer:
// the pieces of one dfrl_trainer. The trainer is created by the first rollout.
struct rollout_store {
  dfrl_env *env = nullptr;
  model *policy_model = nullptr;
  int device_rule = -1;  // DFRL_HEUR_*: a rule-based policy instead of a net
  bool deterministic = false;
  std::size_t obs_cols = 0;
  int algo = -1;  // DFRL_ALGO_*, -1: no learner attached (evaluation only)
  model *value_model = nullptr;
  optimizer *policy_opt = nullptr, *value_opt = nullptr;
  float gamma = 0.99f, lambda = 0.95f;
  int epochs = 4;
  float kl_target = 1e-9f, kl_beta0 = 1.f;
  dfrl_trainer *trainer = nullptr;
  int work = 0;
  bool rolled = false;
  bool learned = false;  // a learn() ran on this trainer: optimizer state lives on the device
  // evaluation-only bookkeeping (agent::play_one_episode with a deterministic policy)
  double eval_reward = 0;
  long long eval_episodes = 0;

  ~rollout_store() { drop(); }
  void drop() {
    if (trainer)
      dfrl_trainer_destroy(trainer);
    trainer = nullptr;
  }
  // optimizer::set_rate (nn.h:592) after the first rollout: the device learner reads the optimizers'
  // current rates before every learn() (momentum / Adam state, the Adam step counter and the KL beta
  // are kept; only a captured CUDA graph is re-captured).
  void sync_rates() {
    if (!trainer || !policy_opt)
      return;
    check(dfrl_trainer_set_rates(trainer, policy_opt->rate(), policy_opt->weight_decay(),
                                 value_opt ? value_opt->rate() : 0.f, value_opt ? value_opt->weight_decay() : 0.f));
  }
  void ensure_trainer(int w) {
    if (trainer && w == work)
      return;
    if (algo < 0)
      throw xeno::error("no learner is attached to this replay buffer");
    if (!env || !policy_model)
      throw xeno::error("no agent is attached to this replay buffer");
    // The record buffers are sized by the rollout length. Re-creating the device learner would silently
    // drop momentum / Adam state, the Adam step counter, the KL beta and the statistics: refuse once
    // anything has been learned (a different length BEFORE the first learn() is harmless).
    if (trainer && learned)
      throw xeno::error("the rollout length changed after learn(): play_steps(n) must keep its n "
                        "(the device records and the optimizer state belong to one rollout length)");
    drop();
    dfrl_trainer_config c;
    dfrl_trainer_config_default(&c);
    c.algo = algo;
    c.work = w;
    c.gamma = gamma;
    c.lambda = lambda;
    c.epochs = epochs;
    c.kl_target = kl_target;
    c.kl_beta0 = kl_beta0;
    c.policy_opt = policy_opt->kind();
    c.policy_lr = policy_opt->rate();
    c.policy_wd = policy_opt->weight_decay();
    c.adam_beta1 = policy_opt->beta1();
    c.adam_beta2 = policy_opt->beta2();
    if (value_opt) {
      c.value_opt = value_opt->kind();
      c.value_lr = value_opt->rate();
      c.value_wd = value_opt->weight_decay();
    }
    c.action_mode = deterministic ? DFRL_ACT_ARGMAX : DFRL_ACT_SAMPLE;
    check(dfrl_trainer_create(device::get(), &c, env, policy_model->handle(obs_cols),
                              value_model ? value_model->handle(obs_cols) : nullptr, &trainer));
    work = w;
  }
};

// transition / trajectory / td (reference rl.h:111-208) with the reference's members. On the device a
// transition is one column of the rollout record; these host objects exist only in what sample_td() returns.
template <typename A, typename S> struct transition {
  transition(A &&a, float r, S &&next) : action(std::move(a)), reward(r), end_state(std::move(next)) {}
  const S *start_state = nullptr;  // the trajectory's opening state or the previous transition's end state
  A action;                        // choice + the policy output it was drawn from (action.distrib)
  float reward = 0.f;
  S end_state;
};
template <typename A, typename S> struct trajectory {
  explicit trajectory(S &&o) : opening(std::move(o)) {}
  const S &last_state() const { return transitions.empty() ? opening : transitions.back().end_state; }
  std::size_t size() const { return transitions.size(); }
  void add_transition(A &&a, float r, S &&next) {
    transitions.emplace_back(std::move(a), r, std::move(next));
  }
  void fill_reference() {  // start_state pointers along the chain (std::list nodes do not move)
    const S *prev = &opening;
    for (auto &x : transitions)
      x.start_state = prev, prev = &x.end_state;
  }
  void freeze() { frozen = true, fill_reference(); }
  S opening;
  std::list<transition<A, S>> transitions;
  bool frozen = false;  // the episode ended inside the record (game_over)
};
// Temporal differences: a read-only window on one trajectory.
template <typename A, typename S> class td {
public:
  using container = std::list<transition<A, S>>;
  explicit td(const trajectory<A, S> &traj) : traj_(&traj) {}
  typename container::const_iterator begin() const { return traj_->transitions.begin(); }
  typename container::const_iterator end() const { return traj_->transitions.end(); }
  std::size_t size() const { return traj_->transitions.size(); }
  bool frozen() const { return traj_->frozen; }
  const transition<A, S> &front() const { return traj_->transitions.front(); }
  const transition<A, S> &back() const { return traj_->transitions.back(); }

private:
  const trajectory<A, S> *traj_;
};
template <typename A, typename S> float total_rewards(const std::vector<td<A, S>> &experience) {
  float sum = 0;
  for (const auto &t : experience)
    for (const auto &x : t)
      sum += x.reward;
  return sum;
}

// How an application's state type S is rebuilt from one column of the device record (2B + 2 int8 planes) and
// what state a finished episode ends in (the device derives it from the start state and the action instead of
// storing it). Applications specialise it next to S (apps/bin_packing/bin_packing.h).
template <typename S> struct record_codec;  // static S decode(const int8_t *planes); static S terminal(const S &start, std::size_t choice);

template <typename A, typename S> class replay_buffer {
public:
  // learner::step() then forget() (ppo_training.cc:63-65): the device records are overwritten by
  // the next rollout; open trajectories continue from the live environment state.
  void forget() { host_.clear(); }

  // replay_buffer::sample_td (rl.h:222-234): every trajectory of the LAST rollout, read back from the device
  // records: per environment slot the recorded steps split where an episode ended (frozen trajectories);
  // an open trajectory ends in the slot's live state. O(transitions) host work and one device read per field:
  // an inspection / custom-learner path, not part of the training loop.
  std::vector<td<A, S>> sample_td() {
    materialise();
    std::vector<td<A, S>> out;
    out.reserve(host_.size());
    for (const auto &t : host_)
      out.emplace_back(t);
    return out;
  }

  dfrl_trainer_stats stats() {
    dfrl_trainer_stats s{};
    if (store_.trainer)
      check(dfrl_trainer_get_stats(store_.trainer, &s));
    return s;
  }
  rollout_store &store() { return store_; }

private:
  template <typename T> std::vector<T> field(int f) {
    std::size_t bytes = 0;
    check(dfrl_trainer_field_size(store_.trainer, f, &bytes));
    std::vector<T> v(bytes / sizeof(T));
    check(dfrl_trainer_read(store_.trainer, f, v.data(), bytes));
    return v;
  }
  void materialise() {
    host_.clear();
    if (!store_.trainer || !store_.rolled)
      throw xeno::error("sample_td() before any rollout (evaluation runs keep totals only: total_rewards(rb))");
    const std::size_t B = A::cardinality(), P = 2 * B + 2;
    const std::vector<uint8_t> act = field<uint8_t>(DFRL_F_REC_ACTION), done = field<uint8_t>(DFRL_F_REC_DONE);
    const std::vector<int8_t> state = field<int8_t>(DFRL_F_REC_STATE);
    const std::vector<float> probs = field<float>(DFRL_F_REC_PROBS);
    const std::vector<int32_t> len = field<int32_t>(DFRL_F_REC_LEN);
    const std::size_t n = len.size(), L = act.size() / n, stride = state.size() / (L * P);
    const bool episodic = store_.algo == DFRL_ALGO_REINFORCE;  // REINFORCE records whole episodes: len[i] steps
    std::vector<int8_t> col(P);
    auto start_of = [&](std::size_t t, std::size_t i) {
      for (std::size_t q = 0; q < P; ++q)
        col[q] = state[(t * P + q) * stride + i];
      return record_codec<S>::decode(col.data());
    };
    for (std::size_t i = 0; i < n; ++i) {
      const std::size_t steps = episodic ? (std::size_t)len[i] : L;
      trajectory<A, S> *open = nullptr;
      for (std::size_t t = 0; t < steps; ++t) {
        const std::size_t k = t * n + i;
        S start = start_of(t, i);
        if (!open) {
          host_.emplace_back(S(start));
          open = &host_.back();
        }
        A a;
        a.choice = act[k];
        vector d({B});
        for (std::size_t q = 0; q < B; ++q)
          d[q] = probs[k * B + q];
        a.distrib = std::move(d);
        if (done[k]) {  // the overflowed state of a finished episode: reward 0, trajectory frozen (rl.h:333-346)
          open->add_transition(std::move(a), 0.f, record_codec<S>::terminal(start, act[k]));
          open->freeze();
          open = nullptr;
        } else if (t + 1 < steps) {
          open->add_transition(std::move(a), 1.f, start_of(t + 1, i));
        } else {  // the record ends here: the slot's live state
          check(dfrl_env_view_one(store_.env, (int)i, col.data()));
          open->add_transition(std::move(a), 1.f, record_codec<S>::decode(col.data()));
        }
      }
      if (open)
        open->fill_reference();
    }
  }
  rollout_store store_;
  std::list<trajectory<A, S>> host_;  // what the last sample_td() returned windows on
};

// total_rewards(rb.sample_td()) of the evaluation loop (ppo_training.cc:75-79).
template <typename A, typename S> float total_rewards(replay_buffer<A, S> &rb) {
  rollout_store &s = rb.store();
  if (s.algo < 0)
    return (float)s.eval_reward;
  return (float)rb.stats().reward_sum;
}

template <typename A, typename S> class agent {
public:
  explicit agent(const policy<A, S> &p, environment<A, S> &env, replay_buffer<A, S> &rb, std::size_t id = 0)
      : id_(id), policy_(p), env_(env), replay_buffer_(rb) {
    rollout_store &s = rb.store();
    s.env = env.device_env();
    s.policy_model = p.backing_model();
    s.device_rule = p.device_rule();
    s.deterministic = p.deterministic();
    s.obs_cols = S::length();
    if (!s.env || (!s.policy_model && s.device_rule < 0))
      throw xeno::error("the agent needs a device environment and a model-backed or rule-based policy (no CPU path)");
  }
  virtual ~agent() = default;

  // One transition of EVERY environment slot. Returns whether any episode is still open (always
  // true for the batched environment: finished slots are reset on the device, rl.h:341-346).
  bool step() {
    play_steps(1);
    return true;
  }

  // agent::play_one_episode (rl.h:351-354) for every slot. With a REINFORCE learner attached this
  // is the rollout of one trainer iteration; with a deterministic policy and no learner it is the
  // evaluation loop of the trainer mains / deep_agent.cc.
  void play_one_episode() {
    rollout_store &s = replay_buffer_.store();
    if (s.algo < 0 && s.device_rule >= 0) {  // firstfit_agent.cc / bestfit_agent.cc / minwaste_agent.cc / random_agent.cc
      double total = 0;
      long long steps = 0;
      check(dfrl_heuristic_play(s.env, s.device_rule, 1, &total, &steps));
      s.eval_reward += total;
      s.eval_episodes += (long long)env_.size();
      return;
    }
    if (s.algo < 0) {
      double mean = 0;
      long long steps = 0;
      check(dfrl_eval_argmax(device::get(), s.env, s.policy_model->handle(s.obs_cols), 1, &mean, &steps));
      s.eval_reward += mean * (double)env_.size();
      s.eval_episodes += (long long)env_.size();
      return;
    }
    if (s.algo != DFRL_ALGO_REINFORCE)
      throw xeno::error("play_one_episode needs the REINFORCE learner (episodic records)");
    verify_rules();
    s.ensure_trainer(1);
    check(dfrl_trainer_rollout(s.trainer, nullptr, nullptr, nullptr));
    s.rolled = true;
  }

  void play_steps(std::size_t n) {
    rollout_store &s = replay_buffer_.store();
    if (s.algo == DFRL_ALGO_REINFORCE)
      throw xeno::error("the REINFORCE learner records whole episodes: use play_one_episode()");
    verify_rules();
    s.ensure_trainer((int)n);
    check(dfrl_trainer_rollout(s.trainer, nullptr, nullptr, nullptr));
    s.rolled = true;
  }

  // Parity runs: teacher-forced rollout from host tapes, all [n][size()] step-major (dfrl.h).
  void play_steps(std::size_t n, const uint8_t *items, const uint8_t *actions, const double *uniforms) {
    rollout_store &s = replay_buffer_.store();
    verify_rules();
    s.ensure_trainer((int)n);
    check(dfrl_trainer_rollout(s.trainer, items, actions, uniforms));
    s.rolled = true;
  }

  std::size_t id() { return id_; }

protected:
  // The rules these two express (bin_packing.h:94-106: done = some bin dimension negative, reward =
  // done ? 0 : 1) run inside the environment kernel, which never calls back into host code. An
  // override expressing a DIFFERENT rule must not be ignored silently: before the first rollout it is
  // evaluated on a live state of slot id_ and on that state with one bin overflowed, and an answer
  // that differs from the device rule is an error.
  virtual bool game_over(const S &state) = 0;
  virtual float get_reward(const S &state1, const S &state2) = 0;

  void verify_rules() {
    if (rules_checked_)
      return;
    rules_checked_ = true;
    const S live = env_.view(id_);  // finished slots are reset on the device: live states are never terminal
    if (game_over(live) || get_reward(live, live) != 1.f)
      throw xeno::error("agent::game_over / get_reward differ from the rule the environment kernel implements "
                        "(bin_packing.h:94-106); the device path cannot honour the override");
    if constexpr (requires(S t) { t.bins[0].first = -1; }) {
      S over = live;
      over.bins[0].first = -1;
      if (!game_over(over) || get_reward(live, over) != 0.f)
        throw xeno::error("agent::game_over / get_reward differ from the rule the environment kernel implements "
                          "(bin_packing.h:94-106); the device path cannot honour the override");
    }
  }
  bool rules_checked_ = false;

  std::size_t id_;
  const policy<A, S> &policy_;
  environment<A, S> &env_;
  replay_buffer<A, S> &replay_buffer_;
};

template <typename A, typename S> class learner {
public:
  explicit learner(replay_buffer<A, S> &rb, model &policy_model, optimizer &policy_optimizer, float gamma = 0.99)
      : replay_buffer_(rb), policy_model_(policy_model), policy_optimizer_(policy_optimizer), gamma_(gamma) {}
  virtual ~learner() = default;

  void step() { learn(); }

  virtual void learn() = 0;

protected:
  void learn_on_device() {
    rollout_store &s = replay_buffer_.store();
    if (!s.trainer || !s.rolled)
      throw xeno::error("learn() before any rollout");
    s.sync_rates();
    check(dfrl_trainer_learn(s.trainer));
    s.learned = true;
  }
  replay_buffer<A, S> &replay_buffer_;
  model &policy_model_;
  optimizer &policy_optimizer_;
  float gamma_;
};

} // namespace xylo

#endif // XYLO_RL_
