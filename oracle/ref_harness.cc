// oracle/ref_harness.cc -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Drives the UNMODIFIED reference (beehover/dependence_free_rl, mounted read-only at
// /root/reference) from a C ABI so that python tests / golden-vector scripts / the
// cpu_baseline leg of bench.py can run the reference's own classes on chosen inputs.
// No reference source is copied: this TU only #includes the reference headers where
// they lie and is linked with the reference's own tensor.cc / logging.cc / thread.cc
// (recipe: oracle/Makefile, outputs only under oracle/_ref/).
//
// Exactly one TU may include nn.h/rl.h/policy_gradient.h (they define non-inline
// functions, reference xylo/nn.h:12-18, 548-586) -- this is that TU.
//
// Only tests/, __graft_entry__.smoke(), tests/golden/make_golden.py and bench.py's
// cpu_baseline / --impl reference legs may load the resulting library.

#include <apps/bin_packing/bin_packing.h>
#include <xeno/sys/thread.h>

#include <chrono>
#include <cstdint>
#include <cstring>
#include <list>
#include <map>
#include <memory>
#include <vector>

namespace {

using A = bp::action;
using S = bp::observation;
using traj_t = xylo::trajectory<A, S>;
using trans_t = xylo::transition<A, S>;

enum layer_kind { L_DENSE = 0, L_CONV1D = 1, L_RELU = 2, L_SOFTMAX = 3, L_SOFTMAX_CE = 4 };

std::unique_ptr<xylo::layer> make_layer(int kind, int in, int out) {
  switch (kind) {
  case L_DENSE:
    return std::make_unique<xylo::full_layer>(in, out);
  case L_CONV1D:
    return std::make_unique<xylo::convolution1d_1_layer>(in, out);
  case L_RELU:
    return std::make_unique<xylo::relu_activation>();
  case L_SOFTMAX:
    return std::make_unique<xylo::softmax_layer>();
  case L_SOFTMAX_CE:
    return std::make_unique<xylo::softmax_cross_entropy_layer>();
  }
  return nullptr;
}

void build_model(xylo::model &m, int n, const int *kinds, const int *ins, const int *outs) {
  for (int i = 0; i < n; ++i)
    m.add_layer(make_layer(kinds[i], ins[i], outs[i]));
}

xylo::vector to_vec(const float *p, std::size_t n) {
  xylo::vector v({n});
  std::copy(p, p + n, v.data());
  return v;
}
xylo::matrix to_mat(const float *p, std::size_t r, std::size_t c) {
  xylo::matrix m({r, c});
  std::copy(p, p + r * c, m.data());
  return m;
}

void set_params(xylo::model &m, const float *p) {
  std::size_t n = m.parameters().size();
  xylo::vector v = to_vec(p, n);
  m.set_parameters(v);
}

// Exposes the protected virtual next_parameters() and records (gradient, new params).
struct opt_log_entry {
  int iter;
  int which; // 0 = policy optimizer, 1 = value optimizer
  std::vector<float> grad;
  std::vector<float> params;
};
std::vector<opt_log_entry> g_opt_log;
int g_iter = 0;
bool g_record = true;

template <class Base> struct rec_opt : Base {
  template <class... Args>
  rec_opt(int which, Args &&... args) : Base(std::forward<Args>(args)...), which_(which) {}
  xylo::vector next_parameters(const xylo::vector &p, const xylo::vector &g, float rate) override {
    xylo::vector r = Base::next_parameters(p, g, rate);
    if (g_record) {
      opt_log_entry e;
      e.iter = g_iter;
      e.which = which_;
      e.grad.assign(g.data(), g.data() + g.size());
      e.params.assign(r.data(), r.data() + r.size());
      g_opt_log.push_back(std::move(e));
    }
    return r;
  }
  int which_;
};

std::unique_ptr<xylo::optimizer> make_opt(int which, int kind, xylo::model &m, float lr, float wd) {
  switch (kind) {
  case 0:
    return std::make_unique<rec_opt<xylo::sgd_optimizer>>(which, m, lr, wd);
  case 1:
    return std::make_unique<rec_opt<xylo::momentum_optimizer>>(which, m, lr);
  case 2:
    return std::make_unique<rec_opt<xylo::adam_optimizer>>(which, m, lr);
  }
  return nullptr;
}

struct probe_agent : bp::agent {
  using bp::agent::agent;
  traj_t *cur() { return curr_traj_; }
};

// One record per agent::step(), in execution order.
struct step_rec {
  int32_t iter, env, t, action, done;
  float reward;
  int32_t sbins[2 * bp::num_bins];
  int32_t sitem[2];
  int32_t ebins[2 * bp::num_bins];
  int32_t eitem[2];
  int32_t item_after[2]; // env item after the step (post-reset when done)
  float p_old[bp::num_bins];
};
// One record per learner row, in the reference's row order.
struct row_rec {
  int32_t iter, env, t; // t = -1 for the trajectory's end-state row
  int32_t frozen;
  float advantage;
};

std::vector<step_rec> g_steps;
std::vector<row_rec> g_rows;
std::map<const trans_t *, std::pair<int, int>> g_tmap; // transition -> (env, t)
std::vector<float> g_adv_capture;

void fill_obs(const S &s, int32_t *bins, int32_t *item) {
  for (std::size_t i = 0; i < bp::num_bins; ++i) {
    bins[2 * i] = s.bins[i].first;
    bins[2 * i + 1] = s.bins[i].second;
  }
  item[0] = s.item.first;
  item[1] = s.item.second;
}

// Steps one agent once and logs it. Returns whether the episode is still open.
bool logged_step(probe_agent &ag, bp::environment &env, int env_id, int t) {
  traj_t *before = ag.cur();
  S s0 = before ? before->last_state() : env.view(0);
  bool open = ag.step();
  traj_t *after = ag.cur();
  traj_t *tr = before ? before : after;
  if (!g_record)
    return open;
  if (!tr)
    throw std::runtime_error("trajectory opened and closed within one step");
  const trans_t &x = tr->transitions.back();
  step_rec r{};
  r.iter = g_iter;
  r.env = env_id;
  r.t = t;
  r.action = int32_t(x.action.choice);
  r.done = open ? 0 : 1;
  r.reward = x.reward;
  fill_obs(s0, r.sbins, r.sitem);
  fill_obs(x.end_state, r.ebins, r.eitem);
  S now = env.view(0);
  r.item_after[0] = now.item.first;
  r.item_after[1] = now.item.second;
  for (std::size_t i = 0; i < bp::num_bins; ++i)
    r.p_old[i] = x.action.distrib ? (*x.action.distrib)[i] : 0.0f;
  g_steps.push_back(r);
  g_tmap[&x] = {env_id, t};
  return open;
}

template <class Base> struct probe_learner : Base {
  using Base::Base;
  void optimize_action(xylo::matrix_view sm, const std::vector<A> &a, xylo::vector_view adv) override {
    g_adv_capture.assign(adv.begin(), adv.end());
    Base::optimize_action(sm, a, adv);
  }
};

// Row order bookkeeping: call right before learner.step().
void log_rows(xylo::replay_buffer<A, S> &rb, bool with_end_rows, const std::vector<float> &adv_or_empty) {
  auto experience = rb.sample_td();
  std::size_t k = 0;
  for (const auto &traj : experience) {
    int env = -1;
    for (const auto &tr : traj) {
      auto it = g_tmap.find(&tr);
      row_rec r{};
      r.iter = g_iter;
      r.env = env = it->second.first;
      r.t = it->second.second;
      r.frozen = traj.frozen();
      r.advantage = k < adv_or_empty.size() ? adv_or_empty[k] : 0.0f;
      g_rows.push_back(r);
      ++k;
    }
    if (with_end_rows) {
      row_rec r{};
      r.iter = g_iter;
      r.env = env;
      r.t = -1;
      r.frozen = traj.frozen();
      r.advantage = k < adv_or_empty.size() ? adv_or_empty[k] : 0.0f;
      g_rows.push_back(r);
      ++k;
    }
  }
}

thread_local std::string g_err;

} // namespace

extern "C" {

const char *ref_last_error() { return g_err.c_str(); }

void ref_seed(unsigned seed) { xylo::default_generator().seed(seed); }

// Raw engine outputs, for pinning the oracle's minstd_rand0 restatement.
void ref_engine_draw(unsigned seed, int n, uint32_t *out) {
  std::default_random_engine e(seed);
  for (int i = 0; i < n; ++i)
    out[i] = uint32_t(e());
}

int ref_num_bins() { return int(bp::num_bins); }

int ref_param_count(int n, const int *kinds, const int *ins, const int *outs) {
  xylo::model m;
  build_model(m, n, kinds, ins, outs);
  return int(m.parameters().size());
}

// Reference initialisation (nn.h:12-18, 68-69, 123-124) drawn from the global engine.
int ref_init_params(unsigned seed, int n, const int *kinds, const int *ins, const int *outs, float *params) {
  ref_seed(seed);
  xylo::model m;
  build_model(m, n, kinds, ins, outs);
  xylo::vector p = m.parameters();
  std::copy(p.data(), p.data() + p.size(), params);
  return int(p.size());
}

// model::eval (nn.h:473-479)
int ref_model_eval(int n, const int *kinds, const int *ins, const int *outs, const float *params,
                   const float *x, int rows, int cols, float *y, int ycols) {
  try {
    xylo::model m;
    build_model(m, n, kinds, ins, outs);
    set_params(m, params);
    xylo::matrix X = to_mat(x, rows, cols);
    xylo::matrix Y = m.eval(X);
    if (int(Y.shape()[1]) != ycols)
      throw std::runtime_error("ycols mismatch");
    std::copy(Y.data(), Y.data() + Y.size(), y);
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
}

// model::forward + model::gradient (nn.h:481-488, 510-528) for a given loss gradient dY at
// the model output. acts (optional) receives every layer output, concatenated row-major.
int ref_model_gradient(int n, const int *kinds, const int *ins, const int *outs, const float *params,
                       const float *x, int rows, int cols, const float *dy, int ycols, float *grad,
                       float *out) {
  try {
    xylo::model m;
    build_model(m, n, kinds, ins, outs);
    set_params(m, params);
    xylo::matrix X = to_mat(x, rows, cols);
    std::vector<xylo::matrix> acts = m.forward(X);
    xylo::matrix output = acts.back();
    acts.pop_back();
    if (out)
      std::copy(output.data(), output.data() + output.size(), out);
    xylo::matrix target = to_mat(dy, rows, ycols);
    xylo::vector g = m.gradient(acts, target);
    std::copy(g.data(), g.data() + g.size(), grad);
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
}

// Single layer forward / backward / gradient (nn.h:20-33 interface).
int ref_layer(int kind, int in, int out, const float *params, const float *x, int rows, int xcols,
              const float *dy, int ycols, float *y, float *dx, float *grad) {
  try {
    auto l = make_layer(kind, in, out);
    std::size_t np = l->parameters().size();
    if (np) {
      xylo::vector p = to_vec(params, np);
      l->parameters() = xylo::vector_view(p);
    }
    xylo::matrix X = to_mat(x, rows, xcols);
    xylo::matrix Y = l->forward(X);
    if (int(Y.shape()[1]) != ycols)
      throw std::runtime_error("ycols mismatch");
    if (y)
      std::copy(Y.data(), Y.data() + Y.size(), y);
    if (dy) {
      xylo::matrix dY = to_mat(dy, rows, ycols);
      if (dx) {
        xylo::matrix dX = l->backward(X, dY);
        std::copy(dX.data(), dX.data() + dX.size(), dx);
      }
      if (grad && np) {
        xylo::vector g = l->gradient(X, dY);
        std::copy(g.data(), g.data() + g.size(), grad);
      }
    }
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
}

// discrete_action loss-gradient rules (rl.h:45-74). kind 0 = softmax_gradient_log,
// 1 = clipped_gradient, 2 = gradient_log.
int ref_action_gradient(int kind, const float *p, const float *p_old, int choice, float adv, float *out) {
  try {
    A a;
    a.choice = choice;
    xylo::vector po = to_vec(p_old, bp::num_bins);
    a.distrib = po;
    xylo::vector in = to_vec(p, bp::num_bins);
    xylo::vector o({bp::num_bins});
    if (kind == 0)
      a.softmax_gradient_log(in, o, adv);
    else if (kind == 1)
      a.clipped_gradient(in, o, adv);
    else
      a.gradient_log(in, o, adv);
    std::copy(o.data(), o.data() + o.size(), out);
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
}

// kl_regulated_loss (policy_gradient.h:47-85): rows x num_bins; beta is in/out.
int ref_kl_loss(int rows, const float *p, const float *p_old, const int *choices, const float *adv,
                float d_targ, float *beta, float *out) {
  try {
    std::vector<A> actions(rows);
    for (int i = 0; i < rows; ++i) {
      actions[i].choice = choices[i];
      actions[i].distrib = to_vec(p_old + i * bp::num_bins, bp::num_bins);
    }
    xylo::matrix P = to_mat(p, rows, bp::num_bins);
    xylo::vector a = to_vec(adv, rows);
    xylo::matrix r = xylo::kl_regulated_loss(actions, xylo::vector_view(a), d_targ, *beta, P);
    std::copy(r.data(), r.data() + r.size(), out);
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
}

// std::discrete_distribution on the seeded global engine (tensor.cc:467-470).
void ref_discrete_sample(unsigned seed, const float *w, int n, int count, int32_t *out) {
  ref_seed(seed);
  xylo::vector v = to_vec(w, n);
  for (int i = 0; i < count; ++i)
    out[i] = int32_t(::discrete_distribution(v));
}
int ref_argmax(const float *w, int n) {
  xylo::vector v = to_vec(w, n);
  return int(::argmax(v));
}

// k optimizer updates from k given gradients. kind 0 sgd(wd) / 1 momentum / 2 adam (nn.h:616-698).
int ref_opt_steps(int kind, float lr, float wd, int nparams, const float *params0, int k,
                  const float *grads, float *params_out /* k x nparams */) {
  try {
    xylo::model dummy;
    g_opt_log.clear();
    bool rec = g_record;
    g_record = false;
    auto opt = make_opt(0, kind, dummy, lr, wd);
    struct access : xylo::optimizer {
      using xylo::optimizer::next_parameters;
    };
    xylo::vector p = to_vec(params0, nparams);
    for (int i = 0; i < k; ++i) {
      xylo::vector g = to_vec(grads + std::size_t(i) * nparams, nparams);
      xylo::vector np = (opt.get()->*(&access::next_parameters))(p, g, lr);
      std::copy(np.data(), np.data() + nparams, params_out + std::size_t(i) * nparams);
      p = np;
    }
    g_record = rec;
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
}

// Environment dynamics with forced actions on one bp::environment (bin_packing.h:46-85)
// stepped through bp::agent (rl.h:325-349). Returns the number of steps logged; records are
// fetched with ref_steps_copy().
int ref_env_forced(unsigned seed, int nsteps, const int32_t *actions) {
  try {
    ref_seed(seed);
    g_steps.clear();
    g_tmap.clear();
    g_iter = 0;
    struct forced_policy : xylo::policy<A, S> {
      const int32_t *a;
      mutable int k = 0;
      A react(const S &) const override {
        A r;
        r.choice = std::size_t(a[k++]);
        return r;
      }
    } pol;
    pol.a = actions;
    bp::environment env;
    xylo::replay_buffer<A, S> rb;
    probe_agent ag(pol, env, rb);
    for (int t = 0; t < nsteps; ++t)
      logged_step(ag, env, 0, t);
    return int(g_steps.size());
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
}

int ref_step_rec_size() { return int(sizeof(step_rec)); }
int ref_row_rec_size() { return int(sizeof(row_rec)); }
int ref_steps_count() { return int(g_steps.size()); }
int ref_rows_count() { return int(g_rows.size()); }
void ref_steps_copy(void *dst) { std::memcpy(dst, g_steps.data(), g_steps.size() * sizeof(step_rec)); }
void ref_rows_copy(void *dst) { std::memcpy(dst, g_rows.data(), g_rows.size() * sizeof(row_rec)); }
int ref_opt_log_count() { return int(g_opt_log.size()); }
int ref_opt_log_get(int i, int *iter, int *which, float *grad, float *params) {
  if (i < 0 || i >= int(g_opt_log.size()))
    return -1;
  const auto &e = g_opt_log[i];
  *iter = e.iter;
  *which = e.which;
  if (grad)
    std::copy(e.grad.begin(), e.grad.end(), grad);
  if (params)
    std::copy(e.params.begin(), e.params.end(), params);
  return int(e.grad.size());
}

// The reference trainers (pg_training.cc / ac_training.cc / ppo_training.cc / ppo2_training.cc)
// with a finite iteration count. algo: 0 REINFORCE, 1 actor-critic, 2 PPO-clip, 3 PPO-KL.
// `work` = steps per env per iteration (AC/PPO) or episodes per env per iteration (REINFORCE).
// record != 0: agents are stepped sequentially on the calling thread (bit-reproducible) and
//              every step / learner row / optimizer update is logged.
// record == 0: timing mode; rollouts run on `threads` xeno::sys::thread workers exactly like the
//              reference mains (thread per agent when threads >= n_envs), learner single-threaded.
// Returns wall seconds of the timed loop (< 0 on error); *env_steps receives the transitions made.
double ref_train(int algo, unsigned seed, int n_envs, int work, int iters, int threads, int record,
                 int pn, const int *pk, const int *pi, const int *po, const float *pparams,
                 int popt, float plr, float pwd, int vn, const int *vk, const int *vi, const int *vo,
                 const float *vparams, int vopt, float vlr, float vwd, float gamma,
                 float *pparams_out, float *vparams_out, long long *env_steps) {
  try {
    ref_seed(seed);
    g_steps.clear();
    g_rows.clear();
    g_tmap.clear();
    g_opt_log.clear();
    g_record = record != 0;
    g_iter = 0;

    xylo::model action_model;
    build_model(action_model, pn, pk, pi, po);
    if (pparams)
      set_params(action_model, pparams);
    auto action_opt = make_opt(0, popt, action_model, plr, pwd);

    xylo::model value_model;
    std::unique_ptr<xylo::optimizer> value_opt;
    if (algo != 0) {
      build_model(value_model, vn, vk, vi, vo);
      if (vparams)
        set_params(value_model, vparams);
      value_opt = make_opt(1, vopt, value_model, vlr, vwd);
    }

    xylo::replay_buffer<A, S> rb;
    std::vector<bp::environment> envs;
    std::vector<probe_agent> agents;
    envs.reserve(n_envs);
    agents.reserve(n_envs);
    xylo::policy_gradient_policy<A, S> policy(action_model);
    // The engine state after model construction is unknown to callers that pass explicit
    // parameters; re-seed so env construction is reproducible from `seed` alone.
    ref_seed(seed);
    for (int i = 0; i < n_envs; ++i) {
      envs.emplace_back();
      agents.emplace_back(policy, envs[i], rb);
    }
    if (g_record) {
      // Initial item of every env, logged as pseudo-steps with t = -1.
      for (int i = 0; i < n_envs; ++i) {
        step_rec r{};
        r.iter = -1;
        r.env = i;
        r.t = -1;
        S s = envs[i].view(0);
        fill_obs(s, r.sbins, r.sitem);
        fill_obs(s, r.ebins, r.eitem);
        r.item_after[0] = s.item.first;
        r.item_after[1] = s.item.second;
        g_steps.push_back(r);
      }
    }

    std::unique_ptr<xylo::learner<A, S>> learner;
    if (algo == 0)
      learner = std::make_unique<bp::pg_learner>(rb, action_model, *action_opt, gamma);
    else if (algo == 1)
      learner = std::make_unique<probe_learner<bp::ac_learner>>(rb, action_model, *action_opt,
                                                               value_model, *value_opt, gamma);
    else if (algo == 2)
      learner = std::make_unique<probe_learner<bp::ppo_learner>>(rb, action_model, *action_opt,
                                                                value_model, *value_opt, gamma);
    else
      learner = std::make_unique<probe_learner<bp::kl_ppo_learner>>(rb, action_model, *action_opt,
                                                                   value_model, *value_opt, gamma);

    std::list<xeno::sys::thread> pool;
    if (!g_record)
      for (int i = 0; i < threads; ++i)
        pool.emplace_back(xeno::string::strcat("worker", i));

    long long steps_done = 0;
    std::vector<long long> per_thread(threads > 0 ? threads : 1, 0);
    auto t0 = std::chrono::steady_clock::now();
    for (g_iter = 0; g_iter < iters; ++g_iter) {
      if (g_record) {
        for (int i = 0; i < n_envs; ++i) {
          if (algo == 0) {
            int t = 0;
            for (int e = 0; e < work; ++e)
              while (logged_step(agents[i], envs[i], i, t++))
                ;
            steps_done += t;
          } else {
            for (int t = 0; t < work; ++t)
              logged_step(agents[i], envs[i], i, t);
            steps_done += work;
          }
        }
      } else {
        int w = 0;
        for (auto &th : pool) {
          th.run([&, w]() {
            long long n = 0;
            for (int i = w; i < n_envs; i += threads) {
              if (algo == 0) {
                for (int e = 0; e < work; ++e) {
                  while (agents[i].step())
                    ++n;
                  ++n;
                }
              } else {
                agents[i].play_steps(work);
                n += work;
              }
            }
            per_thread[w] += n;
          });
          ++w;
        }
        for (auto &th : pool)
          th.join();
      }

      if (g_record) {
        if (algo == 0) {
          auto *pg = static_cast<bp::pg_learner *>(learner.get());
          auto experience = rb.sample_td();
          xylo::vector adv = pg->get_advantages(experience);
          std::vector<float> a(adv.data(), adv.data() + adv.size());
          log_rows(rb, false, a);
          learner->step();
        } else {
          std::size_t mark = g_rows.size();
          log_rows(rb, true, {});
          learner->step();
          for (std::size_t k = 0; k < g_adv_capture.size() && mark + k < g_rows.size(); ++k)
            g_rows[mark + k].advantage = g_adv_capture[k];
        }
      } else {
        learner->step();
      }
      rb.forget();
      g_tmap.clear();
    }
    auto t1 = std::chrono::steady_clock::now();
    if (!g_record)
      for (auto n : per_thread)
        steps_done += n;
    if (env_steps)
      *env_steps = steps_done;
    if (pparams_out) {
      xylo::vector p = action_model.parameters();
      std::copy(p.data(), p.data() + p.size(), pparams_out);
    }
    if (vparams_out && algo != 0) {
      xylo::vector p = value_model.parameters();
      std::copy(p.data(), p.data() + p.size(), vparams_out);
    }
    g_record = true;
    return std::chrono::duration<double>(t1 - t0).count();
  } catch (const std::exception &e) {
    g_err = e.what();
    g_record = true;
    return -1.0;
  }
}

// deep_agent.cc / heuristic agents: `episodes` episodes with an argmax policy on the given
// model parameters (policy 0) or a heuristic (1 random, 2 first-fit-like "minwaste" from
// minwaste_agent.cc:10-39). Returns mean reward per episode.
double ref_eval_argmax(unsigned seed, int n, const int *kinds, const int *ins, const int *outs,
                       const float *params, int episodes, long long *env_steps) {
  try {
    ref_seed(seed);
    xylo::model m;
    build_model(m, n, kinds, ins, outs);
    set_params(m, params);
    ref_seed(seed);
    bool rec = g_record;
    g_record = false;
    xylo::policy_gradient_deterministic_policy<A, S> policy(m);
    bp::environment env;
    xylo::replay_buffer<A, S> rb;
    bp::agent agent(policy, env, rb);
    for (int i = 0; i < episodes; ++i)
      agent.play_one_episode();
    auto experience = rb.sample_td();
    double total = xylo::total_rewards<A, S>(experience);
    long long steps = 0;
    for (const auto &t : experience)
      steps += t.size();
    if (env_steps)
      *env_steps = steps;
    g_record = rec;
    return total / episodes;
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1.0;
  }
}

} // extern "C"
