// xylo/nn.h -- layer / model / optimizer with the reference's signatures (xylo/nn.h:20-33, 60-194,
// 350-431, 467-542, 589-698), computed on the B200 through the C ABI (include/dfrl.h).
//
// A model's parameters live in HBM inside a dfrl_mlp (flat, reference order [W out x in][b out] per
// parametric layer, nn.h:56-59); parameters() / set_parameters() move the flat vector, which is also
// the checkpoint format of apps/bin_packing/weights.{10,20}. Layers keep a host copy of their slice
// so that layer::parameters() has the reference's meaning.
#ifndef XYLO_NN_
#define XYLO_NN_

#include <cmath>
#include <functional>
#include <memory>
#include <optional>
#include <span>
#include <string>
#include <string_view>
#include <vector>

#include <xylo/tensor.h>

namespace xylo {

// RAII device scratch for host<->device staging of the per-call layer interface.
class device_buffer {
public:
  explicit device_buffer(std::size_t floats) {
    void *p = nullptr;
    check(dfrl_malloc(device::get(), sizeof(float) * (floats ? floats : 1), &p));
    p_ = static_cast<float *>(p);
  }
  device_buffer(const device_buffer &) = delete;
  ~device_buffer() { dfrl_free(device::get(), p_); }
  float *get() const { return p_; }
  void upload(const float *src, std::size_t n) { check(dfrl_memcpy_h2d(device::get(), p_, src, sizeof(float) * n)); }
  void download(float *dst, std::size_t n) const { check(dfrl_memcpy_d2h(device::get(), dst, p_, sizeof(float) * n)); }
  // tensors may already be device resident
  static const float *stage(matrix_view m, std::optional<device_buffer> &tmp) {
    if (m.on_device())
      return m.data();
    tmp.emplace(m.size());
    tmp->upload(m.data(), m.size());
    return tmp->get();
  }

private:
  float *p_;
};

class layer {
public:
  explicit layer(std::string_view name = "") : name_(name) {}
  virtual ~layer() = default;
  virtual matrix forward(matrix_view t) = 0;
  virtual matrix backward(matrix_view input, matrix_view loss) = 0;
  virtual vector gradient(matrix_view input, matrix_view backprop) = 0;
  virtual vector_view parameters() const = 0;

  std::string_view name() { return name_; }

  // what the model hands to dfrl_mlp_create
  virtual int kind() const = 0;
  virtual int in() const { return 0; }
  virtual int out() const { return 0; }

protected:
  std::string name_;
};

// matmul_layer = full_layer (nn.h:60-110): y = x W^T + b.
class matmul_layer : public layer {
public:
  matmul_layer(std::size_t input_size, std::size_t output_size, std::string_view name = "")
      : layer(name), in_(input_size), out_(output_size), parameters_({output_size * input_size + output_size}) {}

  matrix forward(matrix_view t) override {
    std::optional<device_buffer> tx;
    const float *x = device_buffer::stage(t, tx);
    device_buffer p(parameters_.size()), y(t.num_rows() * out_);
    p.upload(parameters_.data(), parameters_.size());
    check(dfrl_dense_forward(device::get(), p.get(), (int)in_, (int)out_, x, (int)t.num_rows(), y.get(), 0));
    matrix r({t.num_rows(), out_});
    y.download(r.data(), r.size());
    return r;
  }
  matrix backward(matrix_view, matrix_view loss) override {
    std::optional<device_buffer> tl;
    const float *dy = device_buffer::stage(loss, tl);
    device_buffer p(parameters_.size()), dx(loss.num_rows() * in_);
    p.upload(parameters_.data(), parameters_.size());
    check(dfrl_dense_backward(device::get(), p.get(), (int)in_, (int)out_, dy, (int)loss.num_rows(), nullptr,
                              dx.get()));
    matrix r({loss.num_rows(), in_});
    dx.download(r.data(), r.size());
    return r;
  }
  vector gradient(matrix_view input, matrix_view backprop) override {
    std::optional<device_buffer> tx, tl;
    const float *x = device_buffer::stage(input, tx), *dy = device_buffer::stage(backprop, tl);
    device_buffer g(parameters_.size());
    check(dfrl_dense_gradient(device::get(), (int)in_, (int)out_, x, dy, (int)input.num_rows(), g.get(), 0));
    vector r({parameters_.size()});
    g.download(r.data(), r.size());
    return r;
  }
  vector_view parameters() const override { return parameters_; }
  int kind() const override { return DFRL_LAYER_DENSE; }
  int in() const override { return (int)in_; }
  int out() const override { return (int)out_; }

protected:
  std::size_t in_, out_;
  vector parameters_;
};
using full_layer = matmul_layer;

// convolution1d_1_layer (nn.h:113-194): the same dense math on (batch * points, channels).
class convolution1d_1_layer : public matmul_layer {
public:
  convolution1d_1_layer(std::size_t in_channels, std::size_t out_channels, std::string_view name = "")
      : matmul_layer(in_channels, out_channels, name) {}
  matrix forward(matrix_view t) override { return unfold(matmul_layer::forward(fold(t, in_)), t.num_rows()); }
  matrix backward(matrix_view input, matrix_view loss) override {
    return unfold(matmul_layer::backward(fold(input, in_), fold(loss, out_)), loss.num_rows());
  }
  vector gradient(matrix_view input, matrix_view backprop) override {
    return matmul_layer::gradient(fold(input, in_), fold(backprop, out_));
  }
  int kind() const override { return DFRL_LAYER_CONV1D_1; }

private:
  static matrix_view fold(matrix_view m, std::size_t ch) {
    if (m.num_cols() % ch)
      throw xeno::error("row width is not a multiple of the channel count");
    return matrix_view(m.data(), {m.size() / ch, ch}, m.on_device());
  }
  static matrix unfold(matrix m, std::size_t rows) {
    matrix r({rows, m.size() / rows});
    std::memcpy(r.data(), m.data(), sizeof(float) * m.size());
    return r;
  }
};

class relu_activation : public layer {
public:
  matrix forward(matrix_view t) override {
    std::optional<device_buffer> tx;
    const float *x = device_buffer::stage(t, tx);
    device_buffer y(t.size());
    check(dfrl_relu_forward(device::get(), x, t.size(), y.get()));
    matrix r(t.shape());
    y.download(r.data(), r.size());
    return r;
  }
  matrix backward(matrix_view input, matrix_view loss) override {
    std::optional<device_buffer> tx, tl;
    const float *x = device_buffer::stage(input, tx), *dy = device_buffer::stage(loss, tl);
    device_buffer dx(input.size());
    check(dfrl_relu_backward(device::get(), x, dy, input.size(), dx.get()));
    matrix r(input.shape());
    dx.download(r.data(), r.size());
    return r;
  }
  vector gradient(matrix_view, matrix_view) override { return vector({0}); }
  vector_view parameters() const override { return vector_view(nullptr, {0}); }
  int kind() const override { return DFRL_LAYER_RELU; }
};

// softmax_layer (nn.h:379-422): exp(x) / sum exp(x) without max subtraction, full Jacobian backward.
class softmax_layer : public layer {
public:
  matrix forward(matrix_view t) override {
    std::optional<device_buffer> tx;
    const float *x = device_buffer::stage(t, tx);
    device_buffer y(t.size());
    check(dfrl_softmax_forward(device::get(), x, (int)t.num_rows(), (int)t.num_cols(), y.get()));
    matrix r(t.shape());
    y.download(r.data(), r.size());
    return r;
  }
  matrix backward(matrix_view input, matrix_view loss) override {
    std::optional<device_buffer> tx, tl;
    const float *x = device_buffer::stage(input, tx), *dy = device_buffer::stage(loss, tl);
    device_buffer dx(input.size());
    check(dfrl_softmax_backward(device::get(), x, dy, (int)input.num_rows(), (int)input.num_cols(), dx.get()));
    matrix r(input.shape());
    dx.download(r.data(), r.size());
    return r;
  }
  vector gradient(matrix_view, matrix_view) override { return vector({0}); }
  vector_view parameters() const override { return vector_view(nullptr, {0}); }
  int kind() const override { return DFRL_LAYER_SOFTMAX; }
};

// softmax_cross_entropy_layer (nn.h:424-431): the loss gradient is already wrt the logits.
class softmax_cross_entropy_layer : public softmax_layer {
public:
  matrix backward(matrix_view, matrix_view loss) override { return matrix(loss); }
  int kind() const override { return DFRL_LAYER_SOFTMAX_CE; }
};

class model {
public:
  model() = default;
  model(const model &) = delete;
  ~model() {
    if (mlp_)
      dfrl_mlp_destroy(mlp_);
  }

  void add_layer(std::unique_ptr<layer> &&l) {
    if (mlp_)
      throw xeno::error("layers cannot be added once the model is on the device");
    layers_.emplace_back(std::move(l));
  }

  // model::eval (nn.h:473-479)
  matrix eval(matrix_view batch) const {
    dfrl_mlp *m = handle(batch.num_cols());
    std::optional<device_buffer> tx;
    const float *x = device_buffer::stage(batch, tx);
    const std::size_t oc = dfrl_mlp_output_cols(m);
    device_buffer y(batch.num_rows() * oc);
    check(dfrl_mlp_eval(m, x, (int)batch.num_rows(), y.get()));
    matrix r({batch.num_rows(), oc});
    y.download(r.data(), r.size());
    return r;
  }

  // model::forward (nn.h:481-488): the input and every layer's output.
  std::vector<matrix> forward(matrix_view batch) const {
    std::vector<matrix> acts;
    acts.emplace_back(batch);
    for (const auto &l : layers_)
      acts.emplace_back(l_forward(*l, acts.back()));
    return acts;
  }

  void set_parameters(vector_view parameters) {
    dfrl_mlp *m = handle(0);
    if ((int)parameters.size() != dfrl_mlp_param_count(m))
      throw xeno::error("parameter count mismatch");
    if (parameters.on_device())
      throw xeno::error("set_parameters takes a host vector");
    check(dfrl_mlp_set_params(m, parameters.data(), (int)parameters.size()));
  }

  vector parameters() {
    dfrl_mlp *m = handle(0);
    vector r({(std::size_t)dfrl_mlp_param_count(m)});
    check(dfrl_mlp_get_params(m, r.data(), (int)r.size()));
    std::size_t off = 0;  // refresh the layers' host slices
    for (auto &l : layers_) {
      vector_view p = l->parameters();
      if (p.size())
        std::memcpy(p.data(), r.data() + off, sizeof(float) * p.size());
      off += p.size();
    }
    return r;
  }

  // model::gradient (nn.h:510-528): flat gradient (SUM over rows) for a loss gradient at the output.
  vector gradient(const std::vector<matrix> &input, const matrix &target) const {
    dfrl_mlp *m = handle(input.at(0).num_cols());
    device_buffer x(input[0].size()), dy(target.size()), g(dfrl_mlp_param_count(m));
    x.upload(input[0].data(), input[0].size());
    dy.upload(target.data(), target.size());
    check(dfrl_mlp_forward_gradient(m, x.get(), (int)input[0].num_rows(), dy.get(), g.get(), nullptr));
    vector r({(std::size_t)dfrl_mlp_param_count(m)});
    g.download(r.data(), r.size());
    return r;
  }

  std::span<std::unique_ptr<layer>> layers() { return layers_; }

  // Device handle (created on first use). input_cols is needed when the first layer is a conv1d:
  // its row width is points * channels (observation::length()), which the layers do not know.
  dfrl_mlp *handle(std::size_t input_cols) const {
    if (mlp_)
      return mlp_;
    if (layers_.empty())
      throw xeno::error("empty model");
    std::vector<int> kinds, ins, outs;
    for (const auto &l : layers_) {
      kinds.push_back(l->kind());
      ins.push_back(l->in());
      outs.push_back(l->out());
    }
    int cols = (int)input_cols;
    if (layers_[0]->kind() == DFRL_LAYER_DENSE)
      cols = layers_[0]->in();
    if (cols <= 0)
      throw xeno::error("the input width of a conv1d-first model is unknown: evaluate it once or call "
                        "bind_input_cols()");
    check(dfrl_mlp_create(device::get(), (int)layers_.size(), kinds.data(), ins.data(), outs.data(), cols, &mlp_));
    check(dfrl_mlp_init_params(mlp_, seed_));  // nn.h:12-18: dense N(0, 0.01), conv1d He; biases 0
    return mlp_;
  }
  void bind_input_cols(std::size_t cols) const { handle(cols); }
  void set_init_seed(uint64_t s) { seed_ = s; }

private:
  static matrix l_forward(layer &l, const matrix &in) { return l.forward(in); }
  std::vector<std::unique_ptr<layer>> layers_;
  mutable dfrl_mlp *mlp_ = nullptr;
  uint64_t seed_ = 1;
};

// output -> loss gradient at the output (nn.h:545)
using loss_grad_func = std::function<matrix(matrix_view)>;

// square_loss_grad (nn.h:548-550)
inline matrix square_loss_grad(matrix_view output, matrix_view target) {
  matrix r(output.shape());
  for (std::size_t i = 0; i < r.size(); ++i)
    r.data()[i] = output.data()[i] - target.data()[i];
  return r;
}

class optimizer {
public:
  optimizer(model &m, float rate) : model_(m), rate_(rate) {}
  virtual ~optimizer() = default;
  void set_rate(float rate) { rate_ = rate; }
  float rate() const { return rate_; }
  model &target() { return model_; }

  // optimizer::step (nn.h:594-605): forward, loss gradient (host callback, as in the reference),
  // gradient, in-place device update.
  void step(matrix_view input, const loss_grad_func &loss_grad) {
    dfrl_mlp *m = model_.handle(input.num_cols());
    dfrl_ctx *c = device::get();
    const int n = dfrl_mlp_param_count(m), rows = (int)input.num_rows();
    const std::size_t oc = dfrl_mlp_output_cols(m);
    std::optional<device_buffer> tx;
    const float *x = device_buffer::stage(input, tx);
    device_buffer out(rows * oc), dy(rows * oc), g(n);
    check(dfrl_mlp_eval(m, x, rows, out.get()));
    matrix output({(std::size_t)rows, oc});
    out.download(output.data(), output.size());
    matrix target = loss_grad(output);
    dy.upload(target.data(), target.size());
    check(dfrl_mlp_forward_gradient(m, x, rows, dy.get(), g.get(), nullptr));
    apply(c, m, g.get(), n);
  }

  // what the fused trainer needs to know (dfrl_trainer_config)
  virtual int kind() const = 0;
  virtual float weight_decay() const { return 0.f; }
  virtual float beta1() const { return 0.9f; }
  virtual float beta2() const { return 0.999f; }

protected:
  void apply(dfrl_ctx *c, dfrl_mlp *m, const float *grad_dev, int n) {
    if (kind() != DFRL_OPT_SGD && !state_)
      state_.emplace((kind() == DFRL_OPT_ADAM ? 2 : 1) * (std::size_t)n), check(dfrl_memset(c, state_->get(), 0, sizeof(float) * (kind() == DFRL_OPT_ADAM ? 2 : 1) * n));
    check(dfrl_opt_step(c, kind(), dfrl_mlp_params_dev(m), grad_dev, state_ ? state_->get() : nullptr, n, rate_,
                        weight_decay(), beta1(), beta2(), t_));
    if (kind() == DFRL_OPT_ADAM)
      t_ += 1;  // nn.h:686
  }

private:
  model &model_;
  float rate_;
  float t_ = 1;  // adam step counter, a float starting at 1 (nn.h:693)
  std::optional<device_buffer> state_;
};

class sgd_optimizer : public optimizer {
public:
  sgd_optimizer(model &m, float rate, float weight_decay = 0.0f) : optimizer(m, rate), weight_decay_(weight_decay) {}
  int kind() const override { return DFRL_OPT_SGD; }
  float weight_decay() const override { return weight_decay_; }

protected:
  float weight_decay_;
};

class momentum_optimizer : public optimizer {
public:
  momentum_optimizer(model &m, float rate) : optimizer(m, rate) {}
  int kind() const override { return DFRL_OPT_MOMENTUM; }
};

class adam_optimizer : public optimizer {
public:
  adam_optimizer(model &m, float rate, float beta1 = 0.9, float beta2 = 0.999)
      : optimizer(m, rate), beta1_(beta1), beta2_(beta2) {}
  int kind() const override { return DFRL_OPT_ADAM; }
  float beta1() const override { return beta1_; }
  float beta2() const override { return beta2_; }

private:
  float beta1_, beta2_;
};

} // namespace xylo

#endif // XYLO_NN_
