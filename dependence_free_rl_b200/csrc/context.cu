// context.cu -- device context, memory, timers, NCCL plumbing of libdfrl_b200.so.
//
// Replaces the reference's stubbed device seam (xylo/tensor.cc:38-39 gpu_alloc/gpu_dealloc
// return nullptr) with real device memory, and adds the one collective the sharded loop needs
// (SUM all-reduce of the flat gradient).  NCCL is dlopen'ed so that a
// single-GPU process has no NCCL dependency at all.
#include <dlfcn.h>
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void dfrl_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char *dfrl_last_error(void) { return g_err; }
extern "C" const char *dfrl_version(void) { return "dfrl-b200 0.1 (sm_100a)"; }

// ------------------------------------------------------------------ NCCL via dlopen ---------
namespace {
typedef struct { char internal[128]; } nccl_uid;
typedef int (*fn_get_uid)(nccl_uid *);
typedef int (*fn_init_rank)(void **, int, nccl_uid, int);
typedef int (*fn_allreduce)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef int (*fn_destroy)(void *);
typedef const char *(*fn_errstr)(int);

struct nccl_api {
  void *handle = nullptr;
  fn_get_uid get_uid = nullptr;
  fn_init_rank init_rank = nullptr;
  fn_allreduce allreduce = nullptr;
  fn_destroy destroy = nullptr;
  fn_errstr errstr = nullptr;
} g_nccl;

int load_nccl() {
  if (g_nccl.handle)
    return DFRL_OK;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.handle)
      break;
  }
  if (!g_nccl.handle) {
    dfrl_set_error("cannot dlopen libnccl.so.2: %s", dlerror());
    return DFRL_ERR_NCCL;
  }
  g_nccl.get_uid = (fn_get_uid)dlsym(g_nccl.handle, "ncclGetUniqueId");
  g_nccl.init_rank = (fn_init_rank)dlsym(g_nccl.handle, "ncclCommInitRank");
  g_nccl.allreduce = (fn_allreduce)dlsym(g_nccl.handle, "ncclAllReduce");
  g_nccl.destroy = (fn_destroy)dlsym(g_nccl.handle, "ncclCommDestroy");
  g_nccl.errstr = (fn_errstr)dlsym(g_nccl.handle, "ncclGetErrorString");
  if (!g_nccl.get_uid || !g_nccl.init_rank || !g_nccl.allreduce || !g_nccl.destroy) {
    dfrl_set_error("libnccl.so.2 lacks required symbols");
    return DFRL_ERR_NCCL;
  }
  return DFRL_OK;
}
// ncclDataType_t: ncclFloat32 = 7, ncclFloat64 = 8; ncclRedOp_t: ncclSum = 0
const int kNcclFloat = 7, kNcclDouble = 8, kNcclSum = 0;

int nccl_check(int rc, const char *what) {
  if (rc == 0)
    return DFRL_OK;
  dfrl_set_error("%s failed: %s", what, g_nccl.errstr ? g_nccl.errstr(rc) : "nccl error");
  return DFRL_ERR_NCCL;
}
}  // namespace

extern "C" int dfrl_nccl_unique_id(void *id128_host) {
  DFRL_CHECK(id128_host, "null id buffer");
  DFRL_TRY(load_nccl());
  nccl_uid id;
  DFRL_TRY(nccl_check(g_nccl.get_uid(&id), "ncclGetUniqueId"));
  memcpy(id128_host, &id, 128);
  return DFRL_OK;
}

// ------------------------------------------------------------------ context ------------------
// Everything a context owns; also the error path of dfrl_init (members that were never created are null).
static void release_ctx(dfrl_ctx *ctx) {
  cudaSetDevice(ctx->device);
  if (ctx->stream)
    cudaStreamSynchronize(ctx->stream);
  if (ctx->nccl_comm && g_nccl.destroy)
    g_nccl.destroy(ctx->nccl_comm);
  if (ctx->scratch)
    cudaFree(ctx->scratch);
  if (ctx->umma_ws)
    cudaFree(ctx->umma_ws);
  for (int r = 0; r < ctx->nranks && ctx->p2p.attached; ++r)
    if (r != ctx->rank && ctx->p2p.peer[r])
      cudaIpcCloseMemHandle(ctx->p2p.peer[r]);
  if (ctx->p2p.local)
    cudaFree(ctx->p2p.local);
  if (ctx->ev0)
    cudaEventDestroy(ctx->ev0);
  if (ctx->ev1)
    cudaEventDestroy(ctx->ev1);
  if (ctx->stream)
    cudaStreamDestroy(ctx->stream);
  delete ctx;
}

static int init_resources(dfrl_ctx *ctx, const void *nccl_id) {
  DFRL_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  DFRL_CUDA(cudaEventCreate(&ctx->ev0));
  DFRL_CUDA(cudaEventCreate(&ctx->ev1));
  if (ctx->nranks > 1) {
    DFRL_CHECK(nccl_id, "nccl_id required for nranks > 1");
    DFRL_TRY(load_nccl());
    nccl_uid id;
    memcpy(&id, nccl_id, 128);
    DFRL_TRY(nccl_check(g_nccl.init_rank(&ctx->nccl_comm, ctx->nranks, id, ctx->rank), "ncclCommInitRank"));
  }
  return DFRL_OK;
}

extern "C" int dfrl_init(int device, int nranks, int rank, const void *nccl_id, dfrl_ctx **out) {
  DFRL_CHECK(out, "null out");
  DFRL_CHECK(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank %d / %d", rank, nranks);
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    dfrl_set_error("no CUDA device: %s (libdfrl_b200 has no CPU fallback)",
                   e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    return DFRL_ERR_CUDA;
  }
  DFRL_CHECK(device >= 0 && device < count, "device %d out of range (%d devices)", device, count);
  DFRL_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  DFRL_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    dfrl_set_error("device %d (%s) is sm_%d%d; libdfrl_b200 is built for sm_100a only", device,
                   prop.name, prop.major, prop.minor);
    return DFRL_ERR_CUDA;
  }
  dfrl_ctx *ctx = new dfrl_ctx();
  ctx->device = device;
  ctx->nranks = nranks;
  ctx->rank = rank;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->cc_major = prop.major;
  ctx->cc_minor = prop.minor;
  ctx->hbm_bytes = prop.totalGlobalMem;
  int rc = init_resources(ctx, nccl_id);
  if (rc != DFRL_OK) {  // release whatever was created (dfrl_last_error keeps the reason)
    release_ctx(ctx);
    return rc;
  }
  *out = ctx;
  return DFRL_OK;
}

extern "C" int dfrl_destroy(dfrl_ctx *ctx) {
  if (!ctx)
    return DFRL_OK;
  release_ctx(ctx);
  return DFRL_OK;
}

extern "C" int dfrl_sync(dfrl_ctx *ctx) {
  DFRL_CHECK(ctx, "null ctx");
  DFRL_CUDA(cudaStreamSynchronize(ctx->stream));
  return DFRL_OK;
}

extern "C" void *dfrl_stream(dfrl_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

extern "C" int dfrl_device_info(dfrl_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor,
                                size_t *hbm_bytes) {
  DFRL_CHECK(ctx, "null ctx");
  if (sm_count) *sm_count = ctx->sm_count;
  if (cc_major) *cc_major = ctx->cc_major;
  if (cc_minor) *cc_minor = ctx->cc_minor;
  if (hbm_bytes) *hbm_bytes = ctx->hbm_bytes;
  return DFRL_OK;
}

extern "C" long long dfrl_launch_count(dfrl_ctx *ctx) { return ctx ? ctx->launches : 0; }

int dfrl_scratch(dfrl_ctx *ctx, size_t bytes, void **out) {
  if (bytes > ctx->scratch_bytes) {
    // the old block may still be in use by enqueued kernels
    DFRL_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->scratch)
      DFRL_CUDA(cudaFree(ctx->scratch));
    ctx->scratch = nullptr;
    ctx->scratch_bytes = 0;
    size_t want = round_up(bytes + bytes / 4, 1 << 20);
    DFRL_CUDA(cudaMalloc(&ctx->scratch, want));
    ctx->scratch_bytes = want;
  }
  *out = ctx->scratch;
  return DFRL_OK;
}

int dfrl_umma_workspace(dfrl_ctx *ctx, size_t bytes, void **out) {
  if (bytes > ctx->umma_ws_bytes) {
    DFRL_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->umma_ws)
      DFRL_CUDA(cudaFree(ctx->umma_ws));
    ctx->umma_ws = nullptr;
    ctx->umma_ws_bytes = 0;
    size_t want = round_up(bytes + bytes / 4, 1 << 20);
    DFRL_CUDA(cudaMalloc(&ctx->umma_ws, want));
    ctx->umma_ws_bytes = want;
  }
  *out = ctx->umma_ws;
  return DFRL_OK;
}

// ------------------------------------------------------------------ memory -------------------
extern "C" int dfrl_malloc(dfrl_ctx *ctx, size_t bytes, void **out_dev) {
  DFRL_CHECK(ctx && out_dev, "null argument");
  *out_dev = nullptr;
  if (bytes == 0)
    return DFRL_OK;
  DFRL_CUDA(cudaMalloc(out_dev, bytes));
  return DFRL_OK;
}
extern "C" int dfrl_free(dfrl_ctx *ctx, void *ptr_dev) {
  DFRL_CHECK(ctx, "null ctx");
  if (ptr_dev) {
    DFRL_CUDA(cudaStreamSynchronize(ctx->stream));
    DFRL_CUDA(cudaFree(ptr_dev));
  }
  return DFRL_OK;
}
extern "C" int dfrl_memcpy_h2d(dfrl_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes) {
  DFRL_CHECK(ctx && (bytes == 0 || (dst_dev && src_host)), "null argument");
  if (bytes) {
    DFRL_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    DFRL_CUDA(cudaStreamSynchronize(ctx->stream));  // src may be pageable and short-lived
  }
  return DFRL_OK;
}
extern "C" int dfrl_memcpy_d2h(dfrl_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes) {
  DFRL_CHECK(ctx && (bytes == 0 || (dst_host && src_dev)), "null argument");
  if (bytes) {
    DFRL_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    DFRL_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return DFRL_OK;
}
extern "C" int dfrl_memcpy_d2d(dfrl_ctx *ctx, void *dst_dev, const void *src_dev, size_t bytes) {
  DFRL_CHECK(ctx && (bytes == 0 || (dst_dev && src_dev)), "null argument");
  if (bytes)
    DFRL_CUDA(cudaMemcpyAsync(dst_dev, src_dev, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  return DFRL_OK;
}
extern "C" int dfrl_memset(dfrl_ctx *ctx, void *dst_dev, int byte, size_t bytes) {
  DFRL_CHECK(ctx && (bytes == 0 || dst_dev), "null argument");
  if (bytes)
    DFRL_CUDA(cudaMemsetAsync(dst_dev, byte, bytes, ctx->stream));
  return DFRL_OK;
}

extern "C" int dfrl_malloc_host(dfrl_ctx *ctx, size_t bytes, void **out_host) {
  DFRL_CHECK(ctx && out_host, "null argument");
  *out_host = nullptr;
  if (bytes)
    DFRL_CUDA(cudaMallocHost(out_host, bytes));
  return DFRL_OK;
}
extern "C" int dfrl_free_host(dfrl_ctx *ctx, void *ptr_host) {
  DFRL_CHECK(ctx, "null ctx");
  if (ptr_host)
    DFRL_CUDA(cudaFreeHost(ptr_host));
  return DFRL_OK;
}

// ------------------------------------------------------------------ per-kernel profiling ----
namespace {
struct prof_rec {
  const char *name;
  cudaEvent_t e0, e1;
};
struct prof_state {
  std::vector<prof_rec> recs;
  std::vector<cudaEvent_t> pool;
};
}  // namespace

void dfrl_profile_mark(dfrl_ctx *ctx, const char *name, int end) {
  prof_state *ps = (prof_state *)ctx->prof;
  if (!ps)
    return;
  if (!end) {
    prof_rec r;
    r.name = name;
    for (cudaEvent_t *e : {&r.e0, &r.e1}) {
      if (!ps->pool.empty()) {
        *e = ps->pool.back();
        ps->pool.pop_back();
      } else {
        cudaEventCreate(e);
      }
    }
    cudaEventRecord(r.e0, ctx->stream);
    ps->recs.push_back(r);
  } else if (!ps->recs.empty()) {
    cudaEventRecord(ps->recs.back().e1, ctx->stream);
  }
}

extern "C" int dfrl_profile_enable(dfrl_ctx *ctx, int on) {
  DFRL_CHECK(ctx, "null ctx");
  if (on && !ctx->prof)
    ctx->prof = new prof_state();
  ctx->profiling = on ? 1 : 0;
  return DFRL_OK;
}

extern "C" int dfrl_profile_report(dfrl_ctx *ctx, char *buf, size_t cap) {
  DFRL_CHECK(ctx && buf && cap > 0, "null argument");
  buf[0] = 0;
  prof_state *ps = (prof_state *)ctx->prof;
  if (!ps)
    return DFRL_OK;
  DFRL_CUDA(cudaStreamSynchronize(ctx->stream));
  struct agg { std::string name; long long n; double ms; };
  std::vector<agg> out;
  for (prof_rec &r : ps->recs) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    size_t k = 0;
    for (; k < out.size(); ++k)
      if (out[k].name == r.name)
        break;
    if (k == out.size())
      out.push_back({r.name, 0, 0.0});
    out[k].n += 1;
    out[k].ms += ms;
    ps->pool.push_back(r.e0);
    ps->pool.push_back(r.e1);
  }
  ps->recs.clear();
  size_t off = 0;
  for (agg &a : out) {
    int w = snprintf(buf + off, cap - off, "%s %lld %.6f\n", a.name.c_str(), a.n, a.ms);
    if (w < 0 || (size_t)w >= cap - off)
      break;
    off += (size_t)w;
  }
  return DFRL_OK;
}

extern "C" int dfrl_timer_start(dfrl_ctx *ctx) {
  DFRL_CHECK(ctx, "null ctx");
  DFRL_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  return DFRL_OK;
}
extern "C" int dfrl_timer_stop(dfrl_ctx *ctx, float *ms_out) {
  DFRL_CHECK(ctx && ms_out, "null argument");
  DFRL_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  DFRL_CUDA(cudaEventSynchronize(ctx->ev1));
  DFRL_CUDA(cudaEventElapsedTime(ms_out, ctx->ev0, ctx->ev1));
  return DFRL_OK;
}

// ------------------------------------------------------------------ peer memory --------------
// One node, one process per GPU: every rank cudaMalloc's its exchange buffer, exports a CUDA IPC
// handle, the host side gathers the handles of all ranks (any transport) and every rank maps its
// peers' buffers. The fused kernels then read peer gradients straight over NVLink.
extern "C" int dfrl_p2p_export(dfrl_ctx *ctx, void *handle64_host) {
  DFRL_CHECK(ctx && handle64_host, "null argument");
  DFRL_CHECK(ctx->nranks <= DFRL_P2P_MAX_RANKS, "at most %d ranks per node", DFRL_P2P_MAX_RANKS);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  DFRL_CUDA(cudaSetDevice(ctx->device));
  if (!ctx->p2p.local) {
    const size_t bytes = DFRL_P2P_BYTES;
    DFRL_CUDA(cudaMalloc(&ctx->p2p.local, bytes));
    DFRL_CUDA(cudaMemset(ctx->p2p.local, 0, bytes));
  }
  cudaIpcMemHandle_t h;
  DFRL_CUDA(cudaIpcGetMemHandle(&h, ctx->p2p.local));
  memcpy(handle64_host, &h, 64);
  return DFRL_OK;
}

extern "C" int dfrl_p2p_attach(dfrl_ctx *ctx, const void *all_handles_host) {
  DFRL_CHECK(ctx && all_handles_host, "null argument");
  DFRL_CHECK(ctx->p2p.local, "dfrl_p2p_export must be called first");
  DFRL_CHECK(!ctx->p2p.attached, "peers already attached");
  DFRL_CUDA(cudaSetDevice(ctx->device));
  for (int r = 0; r < ctx->nranks; ++r) {
    if (r == ctx->rank) {
      ctx->p2p.peer[r] = ctx->p2p.local;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char *)all_handles_host + 64 * r, 64);
    void *p = nullptr;
    DFRL_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->p2p.peer[r] = (float *)p;
  }
  ctx->p2p.attached = true;
  return DFRL_OK;
}

extern "C" int dfrl_p2p_attached(dfrl_ctx *ctx) { return ctx && ctx->p2p.attached ? 1 : 0; }

// ------------------------------------------------------------------ collectives --------------
extern "C" int dfrl_allreduce_sum(dfrl_ctx *ctx, float *buf_dev, size_t n) {
  DFRL_CHECK(ctx, "null ctx");
  if (ctx->nranks == 1 || n == 0)
    return DFRL_OK;
  DFRL_CHECK(buf_dev, "null buffer");
  return nccl_check(g_nccl.allreduce(buf_dev, buf_dev, n, kNcclFloat, kNcclSum, ctx->nccl_comm,
                                     ctx->stream), "ncclAllReduce");
}
extern "C" int dfrl_allreduce_sum_f64(dfrl_ctx *ctx, double *buf_dev, size_t n) {
  DFRL_CHECK(ctx, "null ctx");
  if (ctx->nranks == 1 || n == 0)
    return DFRL_OK;
  DFRL_CHECK(buf_dev, "null buffer");
  return nccl_check(g_nccl.allreduce(buf_dev, buf_dev, n, kNcclDouble, kNcclSum, ctx->nccl_comm,
                                     ctx->stream), "ncclAllReduce");
}
extern "C" int dfrl_barrier(dfrl_ctx *ctx) {
  DFRL_CHECK(ctx, "null ctx");
  if (ctx->nranks > 1) {
    void *p;
    DFRL_TRY(dfrl_scratch(ctx, 16, &p));
    DFRL_CUDA(cudaMemsetAsync(p, 0, 4, ctx->stream));
    DFRL_TRY(dfrl_allreduce_sum(ctx, (float *)p, 1));
  }
  DFRL_CUDA(cudaStreamSynchronize(ctx->stream));
  return DFRL_OK;
}
