// xylo/device.h -- the seam the reference left open (memory_blob's on_device bit and the stubbed
// gpu_alloc / gpu_dealloc, xylo/tensor.cc:38-39, 78-102), filled with the B200 library. Host code
// sees only the extern "C" prototypes of include/dfrl.h; there is no CPU fallback: without a usable
// sm_100 device every call below throws xeno::error.
#ifndef XYLO_DEVICE_
#define XYLO_DEVICE_

#include <cstdlib>
#include <string>

#include <dfrl.h>
#include <xeno/exception.h>

namespace xylo {

inline void check(int status, std::source_location where = std::source_location::current()) {
  if (status != DFRL_OK)
    throw xeno::error(std::string("dfrl: ") + dfrl_last_error(), where);
}

// One context per process (one process per GPU). DFRL_DEVICE selects the ordinal; multi-GPU runs
// install their own context (rank / NCCL id) with device::install() before anything else.
class device {
public:
  static dfrl_ctx *get() {
    device &d = instance();
    if (!d.ctx_) {
      const char *e = std::getenv("DFRL_DEVICE");
      check(dfrl_init(e ? std::atoi(e) : 0, 1, 0, nullptr, &d.ctx_));
    }
    return d.ctx_;
  }
  static void install(int ordinal, int nranks, int rank, const void *nccl_id128) {
    device &d = instance();
    if (d.ctx_)
      throw xeno::error("device context already created");
    check(dfrl_init(ordinal, nranks, rank, nccl_id128, &d.ctx_));
  }
  // Multi-GPU, after install(): exchange buffers over NVLink peer memory. export_peer_handle()
  // gives this rank's 64-byte CUDA IPC handle; gather the handles of all ranks (any transport) in
  // rank order and attach them. The fused learners then exchange gradients without NCCL.
  static void export_peer_handle(void *handle64) { check(dfrl_p2p_export(get(), handle64)); }
  static void attach_peers(const void *all_handles) { check(dfrl_p2p_attach(get(), all_handles)); }
  static void sync() { check(dfrl_sync(get())); }

private:
  static device &instance() {
    static device d;
    return d;
  }
  device() = default;
  ~device() {
    if (ctx_)
      dfrl_destroy(ctx_);
  }
  dfrl_ctx *ctx_ = nullptr;
};

} // namespace xylo

#endif // XYLO_DEVICE_
