// apps/bin_packing/bin_packing.h -- the bin-packing problem of the reference
// (apps/bin_packing/bin_packing.h:12-149) on the batched device environment.
#ifndef BIN_PACKING
#define BIN_PACKING

#include <sstream>
#include <string>
#include <utility>
#include <vector>

#include <xylo/nn.h>
#include <xylo/policy_gradient.h>

namespace bp {

constexpr std::size_t num_bins = 8;

using action = xylo::discrete_action<num_bins>;

struct observation {
  static std::size_t length() { return 4 * num_bins; }

  static constexpr std::pair<int, int> capacity{8, 8};

  observation(const std::pair<int, int> &bin_shape) : bins(num_bins, bin_shape), item{0, 0} {}

  std::string to_string() const {
    std::ostringstream oss;
    oss << "item: (" << item.first << ", " << item.second << "); bins:";
    for (const auto &b : bins)
      oss << " (" << b.first << ", " << b.second << ")";
    return oss.str();
  }

  // [bin.w / 8, bin.h / 8, item.w / 8, item.h / 8] per bin (bin_packing.h:31-40)
  void to_vector(xylo::vector_view o) const {
    for (std::size_t i = 0; i < bins.size(); ++i) {
      o[4 * i + 0] = float(bins[i].first) / capacity.first;
      o[4 * i + 1] = float(bins[i].second) / capacity.second;
      o[4 * i + 2] = float(item.first) / capacity.first;
      o[4 * i + 3] = float(item.second) / capacity.second;
    }
  }

  std::vector<std::pair<int, int>> bins;
  std::pair<int, int> item;
};

// N independent environments in HBM; `id` selects one. environment() is the reference's single
// environment (bin_packing.h:50-52).
class environment : public xylo::environment<action, observation> {
public:
  static constexpr std::pair<int, int> capacity{8, 8};

  explicit environment(std::size_t n_envs = 1, uint64_t seed = 1234, int64_t env_offset = 0) : n_(n_envs) {
    dfrl_env_config c;
    dfrl_env_config_default(&c);  // 8 bins of (8, 8); items (4, 2) w.p. 0.4 else (1, 2)
    c.n_envs = (int)n_envs;
    c.n_bins = (int)num_bins;
    c.seed = seed;
    c.env_offset = env_offset;
    xylo::check(dfrl_env_create(xylo::device::get(), &c, &env_));
  }
  environment(const environment &) = delete;
  environment(environment &&o) noexcept : n_(o.n_), env_(o.env_) { o.env_ = nullptr; }
  ~environment() override {
    if (env_)
      dfrl_env_destroy(env_);
  }

  void apply(const action &action, std::size_t id) override {
    xylo::check(dfrl_env_apply_one(env_, (int)id, (int)action.choice));
  }
  observation view(std::size_t id) const override {
    int8_t s[2 * num_bins + 2];
    xylo::check(dfrl_env_view_one(env_, (int)id, s));
    observation o(capacity);
    for (std::size_t b = 0; b < num_bins; ++b)
      o.bins[b] = {s[2 * b], s[2 * b + 1]};
    o.item = {s[2 * num_bins], s[2 * num_bins + 1]};
    return o;
  }
  void reset(std::size_t id) override { xylo::check(dfrl_env_reset_one(env_, (int)id)); }

  dfrl_env *device_env() override { return env_; }
  std::size_t size() const override { return n_; }

private:
  std::size_t n_;
  dfrl_env *env_ = nullptr;
};

class agent : public xylo::agent<action, observation> {
public:
  agent(const xylo::policy<action, observation> &p, environment &env, xylo::replay_buffer<action, observation> &rb)
      : xylo::agent<action, observation>(p, env, rb) {}

private:
  bool game_over(const observation &ob) override {
    for (const auto &bin : ob.bins)
      if (bin.first < 0 || bin.second < 0)
        return true;
    return false;
  }
  float get_reward(const observation &, const observation &ob) override { return game_over(ob) ? 0 : 1; }
};

class pg_learner : public xylo::policy_gradient_learner<action, observation> {
public:
  pg_learner(xylo::replay_buffer<action, observation> &rb, xylo::model &action_model,
             xylo::optimizer &action_optimizer, float gamma = 0.99)
      : xylo::policy_gradient_learner<action, observation>(rb, action_model, action_optimizer, gamma) {}
};

class ac_learner : public xylo::actor_critic_learner<action, observation> {
public:
  ac_learner(xylo::replay_buffer<action, observation> &rb, xylo::model &action_model,
             xylo::optimizer &action_optimizer, xylo::model &value_model, xylo::optimizer &value_optimizer,
             float gamma = 0.99)
      : xylo::actor_critic_learner<action, observation>(rb, action_model, action_optimizer, value_model,
                                                        value_optimizer, gamma) {}
};

class ppo_learner : public xylo::ppo_learner<action, observation> {
public:
  ppo_learner(xylo::replay_buffer<action, observation> &rb, xylo::model &action_model,
              xylo::optimizer &action_optimizer, xylo::model &value_model, xylo::optimizer &value_optimizer,
              float gamma = 0.99)
      : xylo::ppo_learner<action, observation>(rb, action_model, action_optimizer, value_model, value_optimizer,
                                               gamma) {}
};

class kl_ppo_learner : public xylo::kl_ppo_learner<action, observation> {
public:
  kl_ppo_learner(xylo::replay_buffer<action, observation> &rb, xylo::model &action_model,
                 xylo::optimizer &action_optimizer, xylo::model &value_model, xylo::optimizer &value_optimizer,
                 float gamma = 0.99)
      : xylo::kl_ppo_learner<action, observation>(rb, action_model, action_optimizer, value_model, value_optimizer,
                                                  gamma) {}
};

} // namespace bp

#endif
