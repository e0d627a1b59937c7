// xylo/tensor.h -- tensor<N> / tensor_view<N> with the reference's surface (xylo/tensor.h:35-319):
// owning tensors, non-owning views, an on_device bit. Host tensors are plain float arrays; device
// tensors (on_device = true) live in HBM through dfrl_malloc and are only touched by kernels --
// element access on them throws, exactly one copy call (to_host / to_device) moves them.
#ifndef XYLO_TENSOR_
#define XYLO_TENSOR_

#include <array>
#include <cstddef>
#include <cstring>
#include <numeric>
#include <span>
#include <vector>

#include <xylo/device.h>

namespace xylo {

template <std::size_t N> using shape_t = std::array<std::size_t, N>;
template <std::size_t N> std::size_t volume(const shape_t<N> &s) {
  std::size_t v = 1;
  for (auto d : s)
    v *= d;
  return v;
}

template <std::size_t N> class tensor;

template <std::size_t N> class tensor_view {
public:
  tensor_view(float *data, const shape_t<N> &shape, bool on_device = false)
      : data_(data), shape_(shape), on_device_(on_device) {}
  tensor_view(const tensor<N> &t);  // views do not own (reference: borrowed blob)

  float *data() const { return data_; }
  const shape_t<N> &shape() const { return shape_; }
  std::size_t size() const { return volume(shape_); }
  bool on_device() const { return on_device_; }
  std::size_t num_rows() const requires(N == 2) { return shape_[0]; }
  std::size_t num_cols() const requires(N == 2) { return shape_[1]; }

  float &operator[](std::size_t i) const requires(N == 1) {
    host_only();
    return data_[i];
  }
  tensor_view<1> operator[](std::size_t i) const requires(N == 2) {
    host_only();
    return tensor_view<1>(data_ + i * shape_[1], {shape_[1]});
  }
  const tensor_view &operator=(float v) const {  // fill (reference: `view = 0`)
    host_only();
    for (std::size_t i = 0; i < size(); ++i)
      data_[i] = v;
    return *this;
  }
  void copy_from(tensor_view<N> o) const {
    if (o.size() != size())
      throw xeno::error("shape mismatch in copy");
    dfrl_ctx *c = device::get();
    if (!on_device_ && !o.on_device_)
      std::memcpy(data_, o.data_, sizeof(float) * size());
    else if (on_device_ && !o.on_device_)
      check(dfrl_memcpy_h2d(c, data_, o.data_, sizeof(float) * size()));
    else if (!on_device_ && o.on_device_)
      check(dfrl_memcpy_d2h(c, data_, o.data_, sizeof(float) * size()));
    else
      check(dfrl_memcpy_d2d(c, data_, o.data_, sizeof(float) * size()));
  }

private:
  void host_only() const {
    if (on_device_)
      throw xeno::error("element access on a device tensor");
  }
  float *data_;
  shape_t<N> shape_;
  bool on_device_;
};

template <std::size_t N> class tensor {
public:
  explicit tensor(const shape_t<N> &shape, bool on_device = false) : shape_(shape), on_device_(on_device) {
    allocate();
  }
  tensor(tensor_view<N> v) : shape_(v.shape()), on_device_(v.on_device()) {
    allocate();
    view().copy_from(v);
  }
  tensor(const tensor &o) : tensor(tensor_view<N>(o)) {}
  tensor(tensor &&o) noexcept : shape_(o.shape_), on_device_(o.on_device_), host_(std::move(o.host_)), dev_(o.dev_) {
    o.dev_ = nullptr;
  }
  tensor &operator=(tensor o) noexcept {
    release();
    shape_ = o.shape_;
    on_device_ = o.on_device_;
    host_ = std::move(o.host_);
    dev_ = o.dev_;
    o.dev_ = nullptr;
    return *this;
  }
  ~tensor() { release(); }

  float *data() { return on_device_ ? dev_ : host_.data(); }
  const float *data() const { return on_device_ ? dev_ : host_.data(); }
  const shape_t<N> &shape() const { return shape_; }
  std::size_t size() const { return volume(shape_); }
  bool on_device() const { return on_device_; }
  std::size_t num_rows() const requires(N == 2) { return shape_[0]; }
  std::size_t num_cols() const requires(N == 2) { return shape_[1]; }
  tensor_view<N> view() const { return tensor_view<N>(const_cast<float *>(data()), shape_, on_device_); }
  decltype(auto) operator[](std::size_t i) const { return view()[i]; }
  tensor &operator=(float v) {
    view() = v;
    return *this;
  }
  tensor to_device() const {
    tensor r(shape_, true);
    r.view().copy_from(view());
    return r;
  }
  tensor to_host() const {
    tensor r(shape_, false);
    r.view().copy_from(view());
    return r;
  }

private:
  void allocate() {
    if (on_device_) {
      void *p = nullptr;
      check(dfrl_malloc(device::get(), sizeof(float) * (size() ? size() : 1), &p));
      dev_ = static_cast<float *>(p);
    } else {
      host_.assign(size(), 0.f);
    }
  }
  void release() {
    if (dev_)
      dfrl_free(device::get(), dev_);
    dev_ = nullptr;
  }
  shape_t<N> shape_;
  bool on_device_;
  std::vector<float> host_;
  float *dev_ = nullptr;
};

template <std::size_t N>
tensor_view<N>::tensor_view(const tensor<N> &t)
    : data_(const_cast<float *>(t.data())), shape_(t.shape()), on_device_(t.on_device()) {}

using vector = tensor<1>;
using matrix = tensor<2>;
using vector_view = tensor_view<1>;
using matrix_view = tensor_view<2>;

// borrow_vector (xylo/tensor.h): a view over caller-owned floats, e.g. an mmap'ed checkpoint.
inline vector_view borrow_vector(std::span<float> s, bool on_device = false) {
  return vector_view(s.data(), {s.size()}, on_device);
}
// argmax, first maximum wins (xylo/tensor.cc:464-466).
inline std::size_t argmax(vector_view v) {
  std::size_t best = 0;
  for (std::size_t i = 1; i < v.size(); ++i)
    if (v[best] < v[i])
      best = i;
  return best;
}

} // namespace xylo

#endif // XYLO_TENSOR_
