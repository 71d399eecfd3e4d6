// api.cu — the extern "C" surface of libcdgpu.so (include/cdgpu.h).  Host-side control only:
// argument checks with the reference's error behaviour, device buffers, kernel launches, result
// copies.  No algorithmic CPU path exists here: without a CUDA device every compute entry fails.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "common.cuh"

#define API extern "C" __attribute__((visibility("default")))

static thread_local char g_err[1024];
int cdgpu_set_error(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}

long long g_cdgpu_launches = 0;
API int cdgpu_launch_count(int64_t *count) {
  if (!count) return cdgpu_set_error(CDGPU_EARG, "null pointer");
  *count = g_cdgpu_launches;
  return CDGPU_OK;
}
API int cdgpu_version(void) { return CDGPU_VERSION; }
API const char *cdgpu_last_error(void) { return g_err; }
API int cdgpu_device_count(int *count) {
  if (!count) return cdgpu_set_error(CDGPU_EARG, "null pointer");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    *count = 0;
    return cdgpu_set_error(CDGPU_ENODEV, "no CUDA device: %s", cudaGetErrorString(e));
  }
  *count = n;
  return CDGPU_OK;
}
API void cdgpu_default_options(cdgpu_options *o) { // CDOptions() utils.jl:14-20
  o->maxIter = 2000;
  o->optTol = 1e-7;
  o->randomize = 1;
  o->warmStart = 1;
  o->numSteps = 50;
  o->seed = 0;
}
API void cdgpu_default_iter_options(cdgpu_iter_options *o) { // IterLassoOptions() utils.jl:32-39
  o->maxIter = 20;
  o->optTol = 1e-2;
  o->initProcedure = CDGPU_INIT_SCREENING;
  o->_pad = 0;
  o->sinit = 5;
  o->sigma_init = 1.0;
  cdgpu_default_options(&o->optionsCD);
}

// CDGPU_TRACE=1: host-side wall-clock trace of the handle constructors (where do create() stalls come from?)
struct Trace {
  bool on;
  std::chrono::steady_clock::time_point t0, last;
  Trace() : on(getenv("CDGPU_TRACE") != nullptr) { t0 = last = std::chrono::steady_clock::now(); }
  void mark(const char *what) {
    if (!on) return;
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[cdgpu trace] %-28s +%8.3f ms (t=%8.3f)\n", what,
            std::chrono::duration<double, std::milli>(now - last).count(),
            std::chrono::duration<double, std::milli>(now - t0).count());
    last = now;
  }
};

// --------------------------------------------------------------- handles --
static int pool_init(int device);
int cd_use_device(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return cdgpu_set_error(CDGPU_ENODEV, "no CUDA device (%s); libcdgpu has no CPU fallback",
                           e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  if (device < 0 || device >= n) return cdgpu_set_error(CDGPU_EARG, "device %d out of range [0,%d)", device, n);
  CUDA_TRY(cudaSetDevice(device));
  CD_TRY(pool_init(device));
  return CDGPU_OK;
}

// Device memory comes from the device's default stream-ordered pool with the release threshold
// lifted, so the multi-GB Gram / staging buffers of consecutive solves are recycled instead of
// being mapped and unmapped by cudaMalloc/cudaFree on every call (tens of ms each at C2 size).
static int pool_init(int device) {
  static bool done[64] = {false};
  if (device < 64 && done[device]) return CDGPU_OK;
  cudaMemPool_t pool;
  CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, device));
  uint64_t thr = UINT64_MAX;
  CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
  if (device < 64) done[device] = true;
  return CDGPU_OK;
}
template <class T>
static int dalloc(T **p, size_t count) {
  *p = nullptr;
  if (count == 0) count = 1;
  CUDA_TRY(cudaMallocAsync((void **)p, count * sizeof(T), (cudaStream_t)0));
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)0)); // usable from any stream from here on
  return CDGPU_OK;
}
static void dfree(void *p) {
  if (p) cudaFreeAsync(p, (cudaStream_t)0);
}

// Streams and events come from a per-device free list: cudaStreamCreate / cudaEventCreate go through
// the resource manager and were seen to stall for up to 300 ms when an NVML client (nvidia-smi, a
// clock sampler) polls the same device, so a handle takes a recycled set and returns it on destroy.
static std::mutex g_res_mu;
static std::vector<StreamSet> g_free_sets[64];
static int g_sm_count[64] = {0};

int stream_set_acquire(int device, StreamSet *out) {
  {
    std::lock_guard<std::mutex> lk(g_res_mu);
    if (device < 64 && !g_free_sets[device].empty()) {
      *out = g_free_sets[device].back();
      g_free_sets[device].pop_back();
      return CDGPU_OK;
    }
  }
  CUDA_TRY(cudaStreamCreateWithFlags(&out->stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreate(&out->ev0));
  CUDA_TRY(cudaEventCreate(&out->ev1));
  return CDGPU_OK;
}
void stream_set_release(int device, const StreamSet &s) {
  std::lock_guard<std::mutex> lk(g_res_mu);
  if (device < 64 && g_free_sets[device].size() < 16) {
    g_free_sets[device].push_back(s);
    return;
  }
  cudaEventDestroy(s.ev0);
  cudaEventDestroy(s.ev1);
  cudaStreamDestroy(s.stream);
}
static int device_sm_count(int device, int *out) {
  if (device < 64 && g_sm_count[device] > 0) {
    *out = g_sm_count[device];
    return CDGPU_OK;
  }
  int v = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device));
  if (device < 64) g_sm_count[device] = v;
  *out = v;
  return CDGPU_OK;
}

// iterate + sweep scratch of a handle: ONE pool allocation carved into 256-byte aligned pieces
static int handle_common_alloc(cdgpu_handle_s *h) {
  const size_t p = (size_t)h->p;
  StreamSet ss;
  CD_TRY(stream_set_acquire(h->device, &ss));
  h->stream = ss.stream;
  h->ev0 = ss.ev0;
  h->ev1 = ss.ev1;
  size_t off = 0;
  auto take = [&off](size_t bytes) {
    const size_t at = off;
    off += (bytes + 255) & ~(size_t)255;
    return at;
  };
  const size_t o_beta = take(p * sizeof(double)), o_act = take(p * sizeof(int)), o_actval = take(p * sizeof(double)),
               o_nact = take(sizeof(int)), o_inlist = take(p), o_omega = take(p * sizeof(double)),
               o_scr = take((cd_scr_tail(p, (size_t)h->n) + 16 * p + 8 + 32) * sizeof(double)),
               o_iscr = take((10 * p + 64 + 4 * 512 + 32) * sizeof(int)), o_bscr = take(4 * p + 64), o_flag = take(16 * sizeof(int)),
               o_chain = take(CD_MULTI_SCR_BYTES), o_rsnap = take(2 * (size_t)(h->n > 0 ? h->n : 0) * sizeof(double) + 16);
  CD_TRY(dalloc(&h->dcommon, off));
  unsigned char *base = h->dcommon;
  h->dbeta = (double *)(base + o_beta);
  h->dact = (int *)(base + o_act);
  h->dactval = (double *)(base + o_actval);
  h->dnact = (int *)(base + o_nact);
  h->dinlist = base + o_inlist;
  h->domega = (double *)(base + o_omega);
  h->dscr = (double *)(base + o_scr);
  h->discr = (int *)(base + o_iscr);
  h->dbscr = base + o_bscr;
  h->dflag = (int *)(base + o_flag);
  h->dchain = (double *)(base + o_chain);
  h->drsnap = (double *)(base + o_rsnap);
  CUDA_TRY(cudaMemsetAsync(h->dbeta, 0, p * sizeof(double), h->stream));
  CUDA_TRY(cudaMemsetAsync(h->dinlist, 0, p, h->stream));
  CUDA_TRY(cudaMemsetAsync(h->dnact, 0, sizeof(int), h->stream));
  CUDA_TRY(cudaMemsetAsync(h->dflag, 0, 16 * sizeof(int), h->stream));
  CD_TRY(device_sm_count(h->device, &h->sm_count));
  CUDA_TRY(cudaEventCreate(&h->sw_ev0));
  CUDA_TRY(cudaEventCreate(&h->sw_ev1));
  return CDGPU_OK;
}

API int cdgpu_destroy(cdgpu_handle h) {
  if (!h) return CDGPU_OK;
  if (h->tall) {
    cdgpu_destroy(h->tall);
    h->tall = nullptr;
  }
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->ownX) dfree(h->dX);
  if (h->owny) dfree(h->dy);
  if (h->ownw) dfree(h->dw);
  dfree(h->dstate);
  dfree(h->daux);
  dfree(h->dcommon);
  dfree(h->dgram);
  dfree(h->dtiles);
  dfree(h->dstats);
  dfree(h->dlam);
  dfree(h->dcolptr);
  dfree(h->drowval);
  dfree(h->dnzval);
  dfree(h->dtall);
  if (h->own_lzX) dfree(const_cast<double *>(h->lzX));
  if (h->own_lzy) dfree(const_cast<double *>(h->lzy));
  dfree(h->dslot);
  dfree(h->ddiag);
  dfree(h->dgather);
  dfree(h->dresume);
  dfree(h->dbatch);
  delete[] h->hslot;
  if (h->lz_stream2) {
    cudaStreamSynchronize(h->lz_stream2);
    cudaStreamDestroy(h->lz_stream2);
  }
  if (h->lz_spec_ev) cudaEventDestroy(h->lz_spec_ev);
  dfree(h->dgather2);
  dfree(h->dbatch2);
  if (h->sw_ev0) cudaEventDestroy(h->sw_ev0);
  if (h->sw_ev1) cudaEventDestroy(h->sw_ev1);
  if (h->lz_ev0) cudaEventDestroy(h->lz_ev0);
  if (h->lz_ev1) cudaEventDestroy(h->lz_ev1);
  if (h->stream) stream_set_release(h->device, StreamSet{h->stream, h->ev0, h->ev1});
  delete h;
  return CDGPU_OK;
}

struct HandleGuard { // frees a half-built handle on an error return
  cdgpu_handle_s *h;
  ~HandleGuard() {
    if (h) cdgpu_destroy(h);
  }
  cdgpu_handle_s *release() {
    cdgpu_handle_s *t = h;
    h = nullptr;
    return t;
  }
};

static int tall_attach(cdgpu_handle_s *h);
static int naive_finish(cdgpu_handle_s *h) {
  // r = copy(y) (cd_differentiable_function.jl:54); column weights a_k = sum_i [w_i] X_ik^2
  CD_TRY(dalloc(&h->dstate, (size_t)h->n));
  CD_TRY(dalloc(&h->daux, (size_t)h->p));
  CUDA_TRY(cudaMemcpyAsync(h->dstate, h->dy, (size_t)h->n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  CD_TRY(launch_colsq(h, h->dX, h->ld, (int)h->n, (int)h->p, h->kind == CDGPU_LOSS_WLS ? h->dw : nullptr, h->daux,
                      false));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (h->kind != CDGPU_LOSS_SQRT && !getenv("CDGPU_NO_TALL") &&
      (!naive_fits(h->n, h->kind == CDGPU_LOSS_WLS) || getenv("CDGPU_FORCE_TALL")))
    CD_TRY(tall_attach(h));
  return CDGPU_OK;
}

static int naive_check(cdgpu_handle *out, int loss_kind, const void *X, int64_t n, int64_t p, int64_t ldx,
                       const void *y, const void *w) {
  if (!out || !X || !y) return cdgpu_set_error(CDGPU_EARG, "null pointer");
  if (loss_kind != CDGPU_LOSS_LS && loss_kind != CDGPU_LOSS_WLS && loss_kind != CDGPU_LOSS_SQRT)
    return cdgpu_set_error(CDGPU_EARG, "loss_kind must be LS, WLS or SQRT");
  if ((loss_kind == CDGPU_LOSS_WLS) != (w != nullptr))
    return cdgpu_set_error(CDGPU_EARG, "w must be given iff the loss is CDWeightedLSLoss");
  if (n < 1 || p < 1 || ldx < n) return cdgpu_set_error(CDGPU_EDIM, "DimensionMismatch");
  if (n > 0x7fffffff || p > 0x7fffffff) return cdgpu_set_error(CDGPU_EDIM, "n and p must fit in 31 bits");
  return CDGPU_OK;
}

API int cdgpu_naive_create(cdgpu_handle *out, int loss_kind, const double *X, int64_t n, int64_t p, int64_t ldx,
                           const double *y, const double *w, int device) {
  return api_guard([&]() -> int {
  CD_TRY(naive_check(out, loss_kind, X, n, p, ldx, y, w));
  CD_TRY(cd_use_device(device));
  HandleGuard g{new (std::nothrow) cdgpu_handle_s()};
  cdgpu_handle_s *h = g.h;
  if (!h) return cdgpu_set_error(CDGPU_ENOMEM, "out of host memory");
  h->kind = loss_kind;
  h->device = device;
  h->n = n;
  h->p = p;
  h->ld = (n + 1) & ~(int64_t)1; // 16-byte aligned columns on the device
  CD_TRY(handle_common_alloc(h));
  CD_TRY(dalloc(&h->dX, (size_t)h->ld * (size_t)p));
  h->ownX = true;
  CD_TRY(dalloc(&h->dy, (size_t)n));
  h->owny = true;
  if (h->ld != n) CUDA_TRY(cudaMemsetAsync(h->dX, 0, (size_t)h->ld * p * sizeof(double), h->stream));
  CUDA_TRY(cudaMemcpy2DAsync(h->dX, h->ld * sizeof(double), X, ldx * sizeof(double), n * sizeof(double), p,
                             cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaMemcpyAsync(h->dy, y, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  if (w) {
    CD_TRY(dalloc(&h->dw, (size_t)n));
    h->ownw = true;
    CUDA_TRY(cudaMemcpyAsync(h->dw, w, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  }
  CD_TRY(naive_finish(h));
  *out = g.release();
  return CDGPU_OK;
  });
}

API int cdgpu_naive_create_dev(cdgpu_handle *out, int loss_kind, const double *dX, int64_t n, int64_t p, int64_t ldx,
                               const double *dy, const double *dw, int device) {
  return api_guard([&]() -> int {
  CD_TRY(naive_check(out, loss_kind, dX, n, p, ldx, dy, dw));
  CD_TRY(cd_use_device(device));
  HandleGuard g{new (std::nothrow) cdgpu_handle_s()};
  cdgpu_handle_s *h = g.h;
  if (!h) return cdgpu_set_error(CDGPU_ENOMEM, "out of host memory");
  h->kind = loss_kind;
  h->device = device;
  h->n = n;
  h->p = p;
  h->ld = ldx;
  CD_TRY(handle_common_alloc(h));
  h->dX = const_cast<double *>(dX);
  h->dy = const_cast<double *>(dy);
  h->dw = const_cast<double *>(dw);
  CUDA_TRY(cudaDeviceSynchronize()); // the caller's stream may still be producing X / y
  CD_TRY(naive_finish(h));
  *out = g.release();
  return CDGPU_OK;
  });
}

static int quad_finish(cdgpu_handle_s *h, bool check_sym) {
  CD_TRY(dalloc(&h->dstate, (size_t)h->p));
  CD_TRY(dalloc(&h->daux, (size_t)h->p));
  CUDA_TRY(cudaMemsetAsync(h->dstate, 0, (size_t)h->p * sizeof(double), h->stream)); // Ax = zeros(p) :307
  if (check_sym) {
    CUDA_TRY(cudaMemsetAsync(h->dflag, 0, sizeof(int), h->stream));
    CD_TRY(launch_check_symmetric(h, h->dX, h->ld, (int)h->p, h->dflag));
    int bad = 0;
    CUDA_TRY(cudaMemcpyAsync(&bad, h->dflag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (bad) return cdgpu_set_error(CDGPU_EARG, "ArgumentError: A is not symmetric");
  }
  CD_TRY(launch_extract_ainv(h, h->dX, h->ld, (int)h->p, h->daux));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return CDGPU_OK;
}

API int cdgpu_quad_create(cdgpu_handle *out, const double *A, int64_t p, int64_t lda, const double *b, int device) {
  return api_guard([&]() -> int {
  if (!out || !A || !b) return cdgpu_set_error(CDGPU_EARG, "null pointer");
  if (p < 1 || lda < p) return cdgpu_set_error(CDGPU_EDIM, "DimensionMismatch");
  if (p > 0x7fffffff) return cdgpu_set_error(CDGPU_EDIM, "p must fit in 31 bits");
  CD_TRY(cd_use_device(device));
  HandleGuard g{new (std::nothrow) cdgpu_handle_s()};
  cdgpu_handle_s *h = g.h;
  if (!h) return cdgpu_set_error(CDGPU_ENOMEM, "out of host memory");
  h->kind = CDGPU_LOSS_QUAD;
  h->device = device;
  h->n = p;
  h->p = p;
  h->ld = (p + 1) & ~(int64_t)1;
  CD_TRY(handle_common_alloc(h));
  CD_TRY(dalloc(&h->dX, (size_t)h->ld * (size_t)p));
  h->ownX = true;
  CD_TRY(dalloc(&h->dy, (size_t)p));
  h->owny = true;
  if (h->ld != p) CUDA_TRY(cudaMemsetAsync(h->dX, 0, (size_t)h->ld * p * sizeof(double), h->stream));
  CUDA_TRY(cudaMemcpy2DAsync(h->dX, h->ld * sizeof(double), A, lda * sizeof(double), p * sizeof(double), p,
                             cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaMemcpyAsync(h->dy, b, p * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CD_TRY(quad_finish(h, true));
  *out = g.release();
  return CDGPU_OK;
  });
}

API int cdgpu_quad_create_dev(cdgpu_handle *out, const double *dA, int64_t p, int64_t lda, const double *db,
                              int device) {
  return api_guard([&]() -> int {
  if (!out || !dA || !db) return cdgpu_set_error(CDGPU_EARG, "null pointer");
  if (p < 1 || lda < p) return cdgpu_set_error(CDGPU_EDIM, "DimensionMismatch");
  if (p > 0x7fffffff) return cdgpu_set_error(CDGPU_EDIM, "p must fit in 31 bits");
  CD_TRY(cd_use_device(device));
  HandleGuard g{new (std::nothrow) cdgpu_handle_s()};
  cdgpu_handle_s *h = g.h;
  if (!h) return cdgpu_set_error(CDGPU_ENOMEM, "out of host memory");
  h->kind = CDGPU_LOSS_QUAD;
  h->device = device;
  h->n = p;
  h->p = p;
  h->ld = lda;
  CD_TRY(handle_common_alloc(h));
  h->dX = const_cast<double *>(dA);
  h->dy = const_cast<double *>(db);
  CUDA_TRY(cudaDeviceSynchronize());
  CD_TRY(quad_finish(h, true));
  *out = g.release();
  return CDGPU_OK;
  });
}

// A = X'X/n, b = -X'y/n on the device, wrapped as a QUAD handle
static int gram_build(cdgpu_handle *out, const double *dX, int64_t n_local, int64_t n_total, int64_t p, int64_t ldx,
                      const double *dy, int device, cdgpu_comm comm, double *pinned_unused);
int cdgpu_comm_allreduce(cdgpu_comm c, double *buf, size_t count, cudaStream_t s); // nccl_comm.cu

static int gram_build(cdgpu_handle *out, const double *dX, int64_t n_local, int64_t n_total, int64_t p, int64_t ldx,
                      const double *dy, int device, cdgpu_comm comm, double *) {
  Trace tr;
  HandleGuard g{new (std::nothrow) cdgpu_handle_s()};
  cdgpu_handle_s *h = g.h;
  if (!h) return cdgpu_set_error(CDGPU_ENOMEM, "out of host memory");
  h->kind = CDGPU_LOSS_QUAD;
  h->device = device;
  h->n = p;
  h->p = p;
  h->ld = (p + 1) & ~(int64_t)1;
  CD_TRY(handle_common_alloc(h));
  tr.mark("handle_common_alloc");
  // A and b live in ONE allocation so a single allreduce covers both
  const size_t na = (size_t)h->ld * (size_t)p;
  CD_TRY(dalloc(&h->dX, na + (size_t)p));
  tr.mark("alloc G");
  h->ownX = true;
  h->dy = h->dX + na;
  h->owny = false;
  CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
  const bool sharded = comm != nullptr;
  CD_TRY(launch_gram(h, dX, n_local, (int)p, ldx, dy, h->dX, h->dy, (double)n_total, sharded ? 0 : 1));
  if (sharded) {
    CD_TRY(cdgpu_comm_allreduce(comm, h->dX, na + (size_t)p, h->stream));
    CD_TRY(launch_scale_gram(h, h->dX, h->dy, (int)p, (double)n_total));
  }
  CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
  tr.mark("launches");
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  tr.mark("gram kernels done");
  float ms = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->gram_ms = ms;
  CD_TRY(quad_finish(h, false)); // symmetric by construction (mirrored tiles)
  tr.mark("quad_finish");
  *out = g.release();
  return CDGPU_OK;
}

API int cdgpu_gram_create_dev(cdgpu_handle *out, const double *dX, int64_t n, int64_t p, int64_t ldx,
                              const double *dy, int device) {
  return api_guard([&]() -> int {
  if (!out || !dX || !dy) return cdgpu_set_error(CDGPU_EARG, "null pointer");
  if (n < 1 || p < 1 || ldx < n) return cdgpu_set_error(CDGPU_EDIM, "DimensionMismatch");
  if (p > 0x7fffffff) return cdgpu_set_error(CDGPU_EDIM, "p must fit in 31 bits");
  Trace tr;
  CD_TRY(cd_use_device(device));
  CUDA_TRY(cudaDeviceSynchronize());
  tr.mark("use_device + device sync");
  return gram_build(out, dX, n, n, p, ldx, dy, device, nullptr, nullptr);
  });
}

API int cdgpu_gram_create_sharded(cdgpu_handle *out, const double *dX_local, int64_t n_local, int64_t n_total,
                                  int64_t p, int64_t ldx, const double *dy_local, cdgpu_comm comm, int device) {
  return api_guard([&]() -> int {
  if (!out || !dX_local || !dy_local) return cdgpu_set_error(CDGPU_EARG, "null pointer");
  if (n_local < 1 || n_total < n_local || p < 1 || ldx < n_local) return cdgpu_set_error(CDGPU_EDIM, "DimensionMismatch");
  if (p > 0x7fffffff) return cdgpu_set_error(CDGPU_EDIM, "p must fit in 31 bits");
  CD_TRY(cd_use_device(device));
  CUDA_TRY(cudaDeviceSynchronize());
  return gram_build(out, dX_local, n_local, n_total, p, ldx, dy_local, device, comm, nullptr);
  });
}

API int cdgpu_gram_create(cdgpu_handle *out, const double *X, int64_t n, int64_t p, int64_t ldx, const double *y,
                          int device) {
  return api_guard([&]() -> int {
  if (!out || !X || !y) return cdgpu_set_error(CDGPU_EARG, "null pointer");
  if (n < 1 || p < 1 || ldx < n) return cdgpu_set_error(CDGPU_EDIM, "DimensionMismatch");
  if (p > 0x7fffffff) return cdgpu_set_error(CDGPU_EDIM, "p must fit in 31 bits");
  CD_TRY(cd_use_device(device));
  // Host X: the rows are staged in chunks on a copy stream while the SYRK of the previous chunk runs
  // (G accumulates over row chunks, then one scale pass), so the H2D transfer hides behind the DMMA
  // work when the host memory is pinned.  The staging copy is dropped afterwards.
  HandleGuard g{new (std::nothrow) cdgpu_handle_s()};
  cdgpu_handle_s *h = g.h;
  if (!h) return cdgpu_set_error(CDGPU_ENOMEM, "out of host memory");
  h->kind = CDGPU_LOSS_QUAD;
  h->device = device;
  h->n = p;
  h->p = p;
  h->ld = (p + 1) & ~(int64_t)1;
  CD_TRY(handle_common_alloc(h));
  const size_t na = (size_t)h->ld * (size_t)p;
  CD_TRY(dalloc(&h->dX, na + (size_t)p));
  h->ownX = true;
  h->dy = h->dX + na;
  h->owny = false;
  const int64_t ld = (n + 1) & ~(int64_t)1;
  double *dXs = nullptr, *dys = nullptr;
  CD_TRY(dalloc(&dXs, (size_t)ld * (size_t)p));
  int rc = dalloc(&dys, (size_t)n);
  cudaStream_t cs = nullptr;
  const int NCH = 8;
  cudaEvent_t ev[NCH] = {nullptr};
  auto cleanup = [&]() {
    if (cs) {
      cudaStreamSynchronize(cs);
      cudaStreamDestroy(cs);
    }
    if (h->stream) cudaStreamSynchronize(h->stream); // queued SYRK launches may still read the staging buffers
    for (int i = 0; i < NCH; ++i)
      if (ev[i]) cudaEventDestroy(ev[i]);
    dfree(dXs);
    dfree(dys);
  };
#define G_TRY(expr)                                                                                                  \
  do {                                                                                                               \
    cudaError_t _e = (expr);                                                                                         \
    if (_e != cudaSuccess) {                                                                                         \
      cleanup();                                                                                                     \
      return cdgpu_set_error(_e == cudaErrorMemoryAllocation ? CDGPU_ENOMEM : CDGPU_ECUDA, "%s: %s", #expr,          \
                             cudaGetErrorString(_e));                                                                \
    }                                                                                                                \
  } while (0)
  if (rc) {
    cleanup();
    return rc;
  }
  G_TRY(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
  G_TRY(cudaEventRecord(h->ev0, h->stream));
  G_TRY(cudaMemcpyAsync(dys, y, n * sizeof(double), cudaMemcpyHostToDevice, cs));
  if (ld != n) G_TRY(cudaMemsetAsync(dXs, 0, (size_t)ld * p * sizeof(double), cs));
  // Row chunks grow geometrically: the first one is small (its H2D is the only exposed transfer), and since
  // PCIe/C2C moves rows ~3x faster than the DMMA SYRK consumes them, each next chunk may be 3x larger and still
  // land before the previous chunk's SYRK ends.  Few launches also means few read-modify-write passes over G.
  // The last launch divides by n in its epilogue (mode 3), so there is no separate scale pass.
  int64_t bounds[NCH + 1];
  int nch = 0;
  bounds[0] = 0;
  {
    int64_t first = 512, grow = 3;
    if (const char *env = getenv("CDGPU_GRAM_FIRST_CHUNK")) first = std::max<int64_t>(16, atoll(env) & ~15ll);
    if (n < 4096) first = n;
    int64_t r0 = 0, len = first;
    while (r0 < n && nch < NCH - 1) {
      const int64_t r1 = std::min<int64_t>(n, r0 + len);
      bounds[++nch] = r1;
      r0 = r1;
      len *= grow;
    }
    if (r0 < n) bounds[++nch] = n;
  }
  for (int ci = 0; ci < nch; ++ci) {
    const int64_t r0 = bounds[ci], rows = bounds[ci + 1] - r0;
    G_TRY(cudaMemcpy2DAsync(dXs + r0, ld * sizeof(double), X + r0, ldx * sizeof(double), rows * sizeof(double), p,
                            cudaMemcpyHostToDevice, cs));
    G_TRY(cudaEventCreateWithFlags(&ev[ci], cudaEventDisableTiming));
    G_TRY(cudaEventRecord(ev[ci], cs));
    G_TRY(cudaStreamWaitEvent(h->stream, ev[ci], 0));
    const int mode = nch == 1 ? 1 : (ci == 0 ? 0 : (ci == nch - 1 ? 3 : 2));
    rc = launch_gram(h, dXs + r0, rows, (int)p, ld, dys + r0, h->dX, h->dy, (double)n, mode);
    if (rc) {
      cleanup();
      return rc;
    }
  }
  G_TRY(cudaEventRecord(h->ev1, h->stream));
  G_TRY(cudaStreamSynchronize(h->stream));
  float ms = 0.f;
  G_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
#undef G_TRY
  h->gram_ms = ms; // includes the overlapped H2D staging
  cleanup();
  CD_TRY(quad_finish(h, false));
  *out = g.release();
  return CDGPU_OK;
  });
}

// ------------------------------------------------------------ host -> device staging --
// A Julia Matrix{Float64} is ordinary pageable memory: cudaMemcpy from it goes through the driver's own small staging
// buffers on ONE thread (10-20 GB/s).  For large inputs the library stages by itself: two pinned buffers (kept for the
// life of the process), several threads copy a block of columns into one while the DMA of the previous block runs
// from the other.  Pinned / registered sources skip the staging and are copied directly.
struct H2DStager {
  std::mutex mu;
  double *buf[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  size_t cap = 0; // doubles per buffer
};
static H2DStager g_stager;
static bool host_is_pageable(const void *p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    (void)cudaGetLastError();
    return true;
  }
  return at.type == cudaMemoryTypeUnregistered;
}
// dst (device, column pitch ld_dst doubles) <- src (host, pitch ld_src): ncols columns of n doubles, in blocks of
// columns; after_block(c0, c1) is called once the copy of columns [c0, c1) has been ENQUEUED on `cs` (the caller
// orders its consumers behind it with an event).  nblocks_hint: blocks wanted by the caller (0: by size).
template <class F>
static int h2d_columns(double *dst, int64_t ld_dst, const double *src, int64_t ld_src, int64_t n, int64_t ncols,
                       cudaStream_t cs, int nblocks_hint, F after_block) {
  const size_t total = (size_t)n * (size_t)ncols * sizeof(double);
  const bool stage = total >= ((size_t)32 << 20) && host_is_pageable(src) && !getenv("CDGPU_NO_STAGING");
  int nthreads = (int)std::min<unsigned>(16, std::max<unsigned>(1, std::thread::hardware_concurrency() - 1));
  if (const char *env = getenv("CDGPU_H2D_THREADS")) nthreads = std::max(1, atoi(env));
  if (!stage) {
    const int nb = std::max(1, nblocks_hint);
    for (int bi = 0; bi < nb; ++bi) {
      const int64_t c0 = ncols * bi / nb, c1 = ncols * (bi + 1) / nb;
      if (c1 <= c0) continue;
      CUDA_TRY(cudaMemcpy2DAsync(dst + c0 * ld_dst, ld_dst * sizeof(double), src + c0 * ld_src, ld_src * sizeof(double),
                                 (size_t)n * sizeof(double), (size_t)(c1 - c0), cudaMemcpyHostToDevice, cs));
      CD_TRY(after_block(c0, c1));
    }
    return CDGPU_OK;
  }
  std::lock_guard<std::mutex> lk(g_stager.mu);
  const size_t want = ((size_t)64 << 20) / sizeof(double); // 64 MiB per buffer
  if (g_stager.cap < want) {
    for (int i = 0; i < 2; ++i) {
      if (g_stager.buf[i]) cudaFreeHost(g_stager.buf[i]);
      g_stager.buf[i] = nullptr;
      CUDA_TRY(cudaHostAlloc((void **)&g_stager.buf[i], want * sizeof(double), cudaHostAllocPortable));
      if (!g_stager.ev[i]) CUDA_TRY(cudaEventCreateWithFlags(&g_stager.ev[i], cudaEventDisableTiming));
    }
    g_stager.cap = want;
  }
  int64_t cols_per = std::max<int64_t>(1, (int64_t)(g_stager.cap / (size_t)ld_dst));
  if (nblocks_hint > 0) cols_per = std::min<int64_t>(cols_per, std::max<int64_t>(1, (ncols + nblocks_hint - 1) / nblocks_hint));
  bool used[2] = {false, false};
  int which = 0;
  for (int64_t c0 = 0; c0 < ncols; c0 += cols_per, which ^= 1) {
    const int64_t c1 = std::min<int64_t>(ncols, c0 + cols_per), nc = c1 - c0;
    if (used[which]) CUDA_TRY(cudaEventSynchronize(g_stager.ev[which])); // its previous DMA has drained
    double *pb = g_stager.buf[which];
    const bool contiguous = ld_dst == n && ld_src == n; // one run of bytes: each thread copies one large piece (the C
                                                        // library then uses streaming stores: no read-for-ownership)
    auto work = [&](int t) {
      if (contiguous) {
        const size_t bytes = (size_t)nc * (size_t)n * sizeof(double);
        const size_t per = ((bytes + (size_t)nthreads - 1) / (size_t)nthreads + 4095) & ~(size_t)4095;
        const size_t b0 = std::min(bytes, per * (size_t)t), b1 = std::min(bytes, b0 + per);
        if (b1 > b0) memcpy(reinterpret_cast<char *>(pb) + b0, reinterpret_cast<const char *>(src + (size_t)c0 * (size_t)ld_src) + b0, b1 - b0);
        return;
      }
      for (int64_t c = c0 + t; c < c1; c += nthreads) {
        memcpy(pb + (size_t)(c - c0) * (size_t)ld_dst, src + (size_t)c * (size_t)ld_src, (size_t)n * sizeof(double));
        if (ld_dst > n) memset(pb + (size_t)(c - c0) * (size_t)ld_dst + n, 0, (size_t)(ld_dst - n) * sizeof(double));
      }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t) th.emplace_back(work, t);
    work(0);
    for (auto &t : th) t.join();
    CUDA_TRY(cudaMemcpyAsync(dst + c0 * ld_dst, pb, (size_t)nc * (size_t)ld_dst * sizeof(double), cudaMemcpyHostToDevice, cs));
    CUDA_TRY(cudaEventRecord(g_stager.ev[which], cs));
    used[which] = true;
    CD_TRY(after_block(c0, c1));
  }
  for (int i = 0; i < 2; ++i)
    if (used[i]) CUDA_TRY(cudaEventSynchronize(g_stager.ev[i])); // the pinned buffers are free again when we return
  return CDGPU_OK;
}

// ------------------------------------------------------- lazy covariance form --
// (lazy_gram.cu) diag(A) and b up front, columns of A = X'X/n formed on demand in batches of 128 by the DMMA GEMM.
static const int LZ_BATCH = 128;  // columns per background batch (one 128-wide DMMA tile column)
static const int LZ_NARROW = 32;  // columns per BLOCKING batch (128 x 32 tiles: a quarter of the tensor work, same X traffic)

static int lazy_alloc(cdgpu_handle_s *h, int64_t n, int64_t p, int device) {
  h->kind = CDGPU_LOSS_QUAD;
  h->device = device;
  h->n = p;
  h->p = p;
  h->ld = (p + 1) & ~(int64_t)1;
  h->lazy = true;
  h->lz_n = n;
  CD_TRY(handle_common_alloc(h));
  int64_t cap = 4096;
  if (const char *env = getenv("CDGPU_LAZY_CAP")) cap = std::max<int64_t>(LZ_BATCH, atoll(env));
  cap = std::min<int64_t>((p + LZ_BATCH - 1) / LZ_BATCH * LZ_BATCH, cap / LZ_BATCH * LZ_BATCH);
  // keep the cache below ~1/4 of the full matrix' footprint for wide problems and below 8 GiB in any case
  while (cap > LZ_BATCH && (size_t)h->ld * (size_t)cap * sizeof(double) > ((size_t)8 << 30)) cap -= LZ_BATCH;
  h->lz_cap = (int)cap;
  h->lz_used = 0;
  CD_TRY(dalloc(&h->dX, (size_t)h->ld * (size_t)cap));
  h->ownX = true;
  CD_TRY(dalloc(&h->dy, (size_t)p));
  h->owny = true;
  CD_TRY(dalloc(&h->dstate, (size_t)p));
  CD_TRY(dalloc(&h->daux, (size_t)p));
  CD_TRY(dalloc(&h->ddiag, (size_t)p));
  CD_TRY(dalloc(&h->dslot, (size_t)p));
  CD_TRY(dalloc(&h->dbatch, (size_t)LZ_BATCH));
  CD_TRY(dalloc(&h->dresume, (size_t)1));
  h->lz_ldb = (n + 1) & ~(int64_t)1;
  CD_TRY(dalloc(&h->dgather, (size_t)h->lz_ldb * (size_t)LZ_BATCH));
  h->hslot = new int[(size_t)p];
  std::fill(h->hslot, h->hslot + p, -1);
  CUDA_TRY(cudaEventCreate(&h->lz_ev0));
  CUDA_TRY(cudaEventCreate(&h->lz_ev1));
  CUDA_TRY(cudaStreamCreateWithFlags(&h->lz_stream2, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreateWithFlags(&h->lz_spec_ev, cudaEventDisableTiming));
  CD_TRY(dalloc(&h->dgather2, (size_t)h->lz_ldb * (size_t)LZ_BATCH));
  CD_TRY(dalloc(&h->dbatch2, (size_t)LZ_BATCH));
  CUDA_TRY(cudaMemsetAsync(h->dstate, 0, (size_t)p * sizeof(double), h->stream)); // Ax = zeros(p) :307
  CUDA_TRY(cudaMemsetAsync(h->dresume, 0, sizeof(CovResume), h->stream));
  CD_TRY(launch_fill_int(h, h->dslot, (int)p, -1));
  return CDGPU_OK;
}

// form the columns `cols` (distinct, not yet cached, at most the free capacity) : gather -> DMMA GEMM -> cache slots
static int lazy_form(cdgpu_handle_s *h, const std::vector<int> &cols) {
  const int64_t p = h->p, n = h->lz_n;
  for (size_t off = 0; off < cols.size(); off += LZ_BATCH) {
    const int nb = (int)std::min<size_t>(LZ_BATCH, cols.size() - off);
    if (h->lz_used + nb > h->lz_cap) return cdgpu_set_error(CDGPU_ECAP, "lazy covariance cache is full");
    CUDA_TRY(cudaMemcpyAsync(h->dbatch, cols.data() + off, (size_t)nb * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CD_TRY(launch_gather_cols(h->stream, h->lzX, h->lz_ldx, n, h->lzw, h->dbatch, nb, nb, h->dgather, h->lz_ldb, h->dslot, h->lz_used));
    CD_TRY(launch_gemm_tn_split(h->stream, h->sm_count, h->lzX, (int)p, h->lz_ldx, h->dgather, nb, h->lz_ldb, n,
                                h->dX + (size_t)h->lz_used * (size_t)h->ld, h->ld, (double)n));
    for (int q = 0; q < nb; ++q) h->hslot[cols[off + (size_t)q]] = h->lz_used + q;
    h->lz_used += nb;
    h->lz_batches += 1;
  }
  return CDGPU_OK;
}

// The cache starts at 4096 slots (CDGPU_LAZY_CAP) and grows by doubling, up to all p columns, when a solve needs more:
// a new buffer, one device-to-device copy of the filled slots (the kernel is paused / not running; a background batch
// has been committed by the caller), the old buffer back to the pool.  Not enough memory for the next size: the cache
// stays as it is and the caller reports CDGPU_ECAP.
static int lazy_reserve(cdgpu_handle_s *h, int64_t slots) {
  if (slots <= h->lz_cap) return CDGPU_OK;
  const int64_t pmax = (h->p + LZ_BATCH - 1) / LZ_BATCH * LZ_BATCH;
  int64_t cap = std::max<int64_t>(h->lz_cap, LZ_BATCH);
  while (cap < slots) cap *= 2;
  cap = std::min(cap, pmax);
  if (cap <= h->lz_cap) return CDGPU_OK;
  double *nb = nullptr;
  if (dalloc(&nb, (size_t)h->ld * (size_t)cap) != CDGPU_OK) {
    (void)cudaGetLastError();
    return CDGPU_OK;
  }
  CUDA_TRY(cudaMemcpyAsync(nb, h->dX, (size_t)h->ld * (size_t)h->lz_used * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  dfree(h->dX);
  h->dX = nb;
  h->lz_cap = (int)cap;
  return CDGPU_OK;
}

// make sure the columns `need` exist; fill the batch up with the best-scoring candidates (|Ax_j + b_j| / omega_j)
static int lazy_ensure(cdgpu_handle_s *h, const std::vector<int> &need, const double *domega, bool speculate) {
  std::vector<int> cols;
  std::vector<unsigned char> taken;
  for (int k : need)
    if (h->hslot[k] < 0 && std::find(cols.begin(), cols.end(), k) == cols.end()) cols.push_back(k);
  if (cols.empty() && (!speculate || h->lz_used > 0)) return CDGPU_OK; // nothing missing (a fresh handle still forms its first batch)
  // a blocking batch is narrow (the sweep kernel is waiting for it); the wide ones are formed in the background.  A
  // solve that keeps pausing is a dense one: from the third pause on the blocking batch doubles every time, so the
  // number of pauses grows with the logarithm of the columns a solve touches and at most ~2x of them are formed.
  int width = getenv("CDGPU_LAZY_SPEC_OFF") ? LZ_BATCH : LZ_NARROW;
  if (h->lz_pauses > 2) width = (int)std::min<int64_t>((int64_t)width << std::min<int64_t>(h->lz_pauses - 2, 6), 2048);
  if (h->lz_used == 0)
    if (const char *env = getenv("CDGPU_LAZY_FIRST")) width = std::max(LZ_NARROW, atoi(env) / LZ_NARROW * LZ_NARROW); // diagnostics
  int target = (int)((cols.size() + width - 1) / width * width);
  if (target == 0) target = width;
  target = (int)std::min<int64_t>(target, std::max<int64_t>((int64_t)cols.size(), h->p - h->lz_used));
  CD_TRY(lazy_reserve(h, (int64_t)h->lz_used + target)); // grows the cache (up to all p columns) when it has to
  const int free_slots = h->lz_cap - h->lz_used;
  if ((int)cols.size() > free_slots)
    return cdgpu_set_error(CDGPU_ECAP, "active set needs more columns than the lazy covariance cache can hold (%d of %lld)",
                           h->lz_cap, (long long)h->p);
  target = std::min(target, free_slots);
  if (speculate && (int)cols.size() < target) {
    const int64_t p = h->p;
    double *dscore = h->dscr + 3 * (size_t)p; // [3p, 4p): free between launches
    CD_TRY(launch_lazy_score(h, h->dstate, h->dy, domega, h->dslot, (int)p, dscore));
    std::vector<double> sc((size_t)p);
    CUDA_TRY(cudaMemcpyAsync(sc.data(), dscore, (size_t)p * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    for (int k : cols) sc[(size_t)k] = -1.0;
    const int want = target - (int)cols.size();
    std::vector<int> idx((size_t)p);
    for (int64_t j = 0; j < p; ++j) idx[(size_t)j] = (int)j;
    const int take = (int)std::min<int64_t>(want, p);
    // the next-best candidates are formed speculatively while the sweep kernel runs (lazy_spec_start): one selection of
    // the take + more best (a strict total order: score, then index), linear in p, then a sort of those few
    const int more = (int)std::min<int64_t>(LZ_BATCH, p - take);
    auto better = [&](int a_, int b_) { return sc[(size_t)a_] > sc[(size_t)b_] || (sc[(size_t)a_] == sc[(size_t)b_] && a_ < b_); };
    if (take + more < p) std::nth_element(idx.begin(), idx.begin() + (take + more), idx.end(), better);
    std::sort(idx.begin(), idx.begin() + (take + more), better);
    for (int q = 0; q < take; ++q)
      if (sc[(size_t)idx[(size_t)q]] > 0.0) cols.push_back(idx[(size_t)q]);
    h->next_n = 0;
    if (more > 0) {
      for (int q = 0; q < more; ++q)
        if (sc[(size_t)idx[(size_t)(take + q)]] > 0.0) h->next_cand[h->next_n++] = idx[(size_t)(take + q)];
    }
  }
  if (cols.empty()) return CDGPU_OK;
  CUDA_TRY(cudaEventRecord(h->lz_ev0, h->stream));
  CD_TRY(lazy_form(h, cols));
  CUDA_TRY(cudaEventRecord(h->lz_ev1, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  float ms = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, h->lz_ev0, h->lz_ev1));
  h->lz_form_ms += ms;
  h->lz_form_cols += (long long)cols.size();
  return CDGPU_OK;
}

// Speculative batch: the runner-up candidates of the last scoring are formed on a second stream while the sweep kernel
// (one 16-CTA cluster) runs on the main stream; the GEMM is sized for the remaining SMs.  The columns become visible to
// the kernel only when the batch is committed (slot map written) at the next pause or at the end of the solve.
static int lazy_spec_start(cdgpu_handle_s *h) {
  if (h->spec_inflight || h->next_n == 0 || getenv("CDGPU_LAZY_SPEC_OFF")) return CDGPU_OK;
  int nb = 0;
  for (int q = 0; q < h->next_n && nb < LZ_BATCH; ++q)
    if (h->hslot[h->next_cand[q]] < 0) h->spec_cols[nb++] = h->next_cand[q];
  h->next_n = 0;
  if (nb == 0 || h->lz_used + nb > h->lz_cap) return CDGPU_OK;
  CUDA_TRY(cudaMemcpyAsync(h->dbatch2, h->spec_cols, (size_t)nb * sizeof(int), cudaMemcpyHostToDevice, h->lz_stream2));
  CD_TRY(launch_gather_cols(h->lz_stream2, h->lzX, h->lz_ldx, h->lz_n, h->lzw, h->dbatch2, nb, nb, h->dgather2, h->lz_ldb, nullptr, 0));
  int spec_sms = std::max(8, h->sm_count - 16);
  if (const char *env = getenv("CDGPU_LAZY_SPEC_SMS")) spec_sms = std::max(8, std::min(spec_sms, atoi(env))); // diagnostics
  CD_TRY(launch_gemm_tn_split(h->lz_stream2, spec_sms, h->lzX, (int)h->p, h->lz_ldx, h->dgather2, nb, h->lz_ldb,
                              h->lz_n, h->dX + (size_t)h->lz_used * (size_t)h->ld, h->ld, (double)h->lz_n));
  CUDA_TRY(cudaEventRecord(h->lz_spec_ev, h->lz_stream2));
  h->spec_n = nb;
  h->spec_slot0 = h->lz_used;
  h->lz_used += nb; // the slots are reserved from now on
  h->spec_inflight = true;
  h->lz_batches += 1;
  return CDGPU_OK;
}
static int lazy_spec_commit(cdgpu_handle_s *h) {
  if (!h->spec_inflight) return CDGPU_OK;
  CUDA_TRY(cudaStreamWaitEvent(h->stream, h->lz_spec_ev, 0));
  CD_TRY(launch_assign_slots(h->stream, h->dbatch2, h->spec_n, h->spec_slot0, h->dslot));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  for (int q = 0; q < h->spec_n; ++q) h->hslot[h->spec_cols[q]] = h->spec_slot0 + q;
  h->spec_inflight = false;
  return CDGPU_OK;
}

static int lazy_finish(cdgpu_handle_s *h) { // diag, b, 1/diag from the resident data
  CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
  CD_TRY(launch_diag_xty(h, h->lzX, h->lz_n, (int)h->p, h->lz_ldx, h->lzy, h->lzw, (double)h->lz_n, h->ddiag, h->dy, h->daux, 0, 1));
  CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  float ms = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->gram_ms = ms;
  return CDGPU_OK;
}

API int cdgpu_gram_create_lazy_dev(cdgpu_handle *out, const double *dX, int64_t n, int64_t p, int64_t ldx,
                                   const double *dy, int device) {
  return api_guard([&]() -> int {
    if (!out || !dX || !dy) return cdgpu_set_error(CDGPU_EARG, "null pointer");
    if (n < 1 || p < 1 || ldx < n) return cdgpu_set_error(CDGPU_EDIM, "DimensionMismatch");
    if (p > 0x7fffffff) return cdgpu_set_error(CDGPU_EDIM, "p must fit in 31 bits");
    CD_TRY(cd_use_device(device));
    CUDA_TRY(cudaDeviceSynchronize()); // the caller's stream may still be producing X / y
    HandleGuard g{new (std::nothrow) cdgpu_handle_s()};
    cdgpu_handle_s *h = g.h;
    if (!h) return cdgpu_set_error(CDGPU_ENOMEM, "out of host memory");
    CD_TRY(lazy_alloc(h, n, p, device));
    h->lzX = dX;
    h->lzy = dy;
    h->lz_ldx = ldx;
    CD_TRY(lazy_finish(h));
    *out = g.release();
    return CDGPU_OK;
  });
}

API int cdgpu_gram_create_lazy(cdgpu_handle *out, const double *X, int64_t n, int64_t p, int64_t ldx, const double *y,
                               int device) {
  return api_guard([&]() -> int {
    if (!out || !X || !y) return cdgpu_set_error(CDGPU_EARG, "null pointer");
    if (n < 1 || p < 1 || ldx < n) return cdgpu_set_error(CDGPU_EDIM, "DimensionMismatch");
    if (p > 0x7fffffff) return cdgpu_set_error(CDGPU_EDIM, "p must fit in 31 bits");
    CD_TRY(cd_use_device(device));
    HandleGuard g{new (std::nothrow) cdgpu_handle_s()};
    cdgpu_handle_s *h = g.h;
    if (!h) return cdgpu_set_error(CDGPU_ENOMEM, "out of host memory");
    CD_TRY(lazy_alloc(h, n, p, device));
    // the data stay resident (columns are formed from them later): one owned copy, 16-byte aligned columns
    const int64_t ld = (n + 1) & ~(int64_t)1;
    double *dXs = nullptr, *dys = nullptr;
    CD_TRY(dalloc(&dXs, (size_t)ld * (size_t)p));
    h->lzX = dXs;
    h->own_lzX = true;
    CD_TRY(dalloc(&dys, (size_t)n));
    h->lzy = dys;
    h->own_lzy = true;
    h->lz_ldx = ld;
    CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
    if (ld != n) CUDA_TRY(cudaMemsetAsync(dXs, 0, (size_t)ld * p * sizeof(double), h->stream));
    CUDA_TRY(cudaMemcpyAsync(dys, y, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    // column blocks: the diag / X'y pass of a block runs while the next block is still on the wire
    cudaStream_t cs = nullptr;
    cudaEvent_t ev[16] = {nullptr};
    int nev = 0;
    auto cleanup = [&]() {
      if (cs) {
        cudaStreamSynchronize(cs);
        cudaStreamDestroy(cs);
      }
      cudaStreamSynchronize(h->stream);
      for (int i = 0; i < nev; ++i) cudaEventDestroy(ev[i]);
    };
    cudaError_t e = cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking);
    if (e != cudaSuccess) return cdgpu_set_error(CDGPU_ECUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    {
      cudaEvent_t e0;
      cudaEventCreateWithFlags(&e0, cudaEventDisableTiming);
      cudaEventRecord(e0, h->stream);
      cudaStreamWaitEvent(cs, e0, 0); // memset / y copy first
      cudaEventDestroy(e0);
    }
    const int NBLK = (int)std::min<int64_t>(16, std::max<int64_t>(1, p / 512));
    int rc = h2d_columns(dXs, ld, X, ldx, n, p, cs, NBLK, [&](int64_t c0, int64_t c1) -> int {
      if (nev >= 16) cudaEventDestroy(ev[--nev]); // more blocks than events (staging splits by size): recycle the last one
      cudaError_t e2 = cudaEventCreateWithFlags(&ev[nev], cudaEventDisableTiming);
      if (e2 == cudaSuccess) {
        nev += 1;
        e2 = cudaEventRecord(ev[nev - 1], cs);
      }
      if (e2 == cudaSuccess) e2 = cudaStreamWaitEvent(h->stream, ev[nev - 1], 0);
      if (e2 != cudaSuccess) return cdgpu_set_error(CDGPU_ECUDA, "H2D staging: %s", cudaGetErrorString(e2));
      return launch_diag_xty(h, dXs + c0 * ld, n, (int)(c1 - c0), ld, dys, nullptr, (double)n, h->ddiag + c0, h->dy + c0, h->daux + c0, 0, 1);
    });
    if (rc == CDGPU_OK) {
      e = cudaEventRecord(h->ev1, h->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
      if (e != cudaSuccess) rc = cdgpu_set_error(CDGPU_ECUDA, "lazy create: %s", cudaGetErrorString(e));
    }
    cleanup();
    CD_TRY(rc);
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->gram_ms = ms; // includes the H2D staging
    *out = g.release();
    return CDGPU_OK;
  });
}

// ------------------------------------------------------------ tall naive-form problems --
// The residual-form sweep kernel keeps r (and w) in shared memory, which bounds n (~28 000 rows, half with weights).
// The reference's CDLeastSquaresLoss / CDWeightedLSLoss have no such bound (cd_differentiable_function.jl:43-194), so a
// handle on a taller problem carries an inner LAZY covariance handle over the same device-resident X, y (w):
// A = X'[W]X/n, b = -X'[W]y/n, the identical minimiser, columns of A formed only for coordinates that become non-zero.
// Solves run there; f.r = y - X beta is formed afterwards from the active columns.  (CDSqrtLassoLoss has no such form:
// tall sqrt-lasso problems run the residual form with the rows dealt over the grid, tall_sweep.cu.)
static int tall_attach(cdgpu_handle_s *h) {
  cdgpu_handle_s *t = new (std::nothrow) cdgpu_handle_s();
  if (!t) return cdgpu_set_error(CDGPU_ENOMEM, "out of host memory");
  h->tall = t; // owned from here on (cdgpu_destroy(h) releases it, also on an error return below)
  CD_TRY(lazy_alloc(t, h->n, h->p, h->device));
  t->lzX = h->dX;
  t->lzy = h->dy;
  t->lzw = h->kind == CDGPU_LOSS_WLS ? h->dw : nullptr;
  t->lz_ldx = h->ld;
  CD_TRY(lazy_finish(t));
  return CDGPU_OK;
}
// f.r = y - X beta for the inner handle's iterate (initialize!, cd_differentiable_function.jl:59-72)
static int tall_residual(cdgpu_handle_s *h) {
  cdgpu_handle_s *t = h->tall;
  NaiveArgs a = {};
  a.X = h->dX;
  a.ldx = h->ld;
  a.n = (int)h->n;
  a.p = (int)h->p;
  a.y = h->dy;
  a.act = t->dact;
  a.actval = t->dactval;
  a.nact = t->dnact;
  a.r = h->dstate;
  a.beta = h->dbeta;
  a.inlist = h->dinlist;
  CUDA_TRY(cudaStreamSynchronize(t->stream));
  CD_TRY(launch_naive_init(h, a));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return CDGPU_OK;
}

API int cdgpu_sweep_ms(cdgpu_handle h, double *ms) {
  if (!h || !ms) return cdgpu_set_error(CDGPU_EARG, "null pointer");
  if (h->sweep_pending) {
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaEventSynchronize(h->sw_ev1));
    float t = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&t, h->sw_ev0, h->sw_ev1));
    h->sweep_ms = t;
    h->sweep_pending = false;
  }
  *ms = h->sweep_ms;
  return CDGPU_OK;
}

API int cdgpu_lazy_stats(cdgpu_handle h, int64_t *columns, int64_t *batches, int64_t *pauses, double *form_ms) {
  if (!h) return cdgpu_set_error(CDGPU_EARG, "null handle");
  if (columns) *columns = h->lz_used;
  if (batches) *batches = h->lz_batches;
  if (pauses) *pauses = h->lz_pauses;
  if (form_ms) *form_ms = h->lz_form_ms;
  return CDGPU_OK;
}

API int cdgpu_lazy_form_columns(cdgpu_handle h, int64_t *columns) {
  if (!h || !columns) return cdgpu_set_error(CDGPU_EARG, "null pointer");
  *columns = h->lz_form_cols;
  return CDGPU_OK;
}

API int cdgpu_dims(cdgpu_handle h, int64_t *n, int64_t *p, int *loss_kind) {
  if (!h) return cdgpu_set_error(CDGPU_EARG, "null handle");
  if (n) *n = h->n;
  if (p) *p = h->p;
  if (loss_kind) *loss_kind = h->kind;
  return CDGPU_OK;
}
API int cdgpu_gram_ms(cdgpu_handle h, double *ms) {
  if (!h || !ms) return cdgpu_set_error(CDGPU_EARG, "null pointer");
  *ms = h->gram_ms;
  return CDGPU_OK;
}
API int cdgpu_quad_get(cdgpu_handle h, double *A_out, double *b_out) {
  return api_guard([&]() -> int {
  if (!h || h->kind != CDGPU_LOSS_QUAD) return cdgpu_set_error(CDGPU_EARG, "not a CDQuadraticLoss handle");
  CUDA_TRY(cudaSetDevice(h->device));
  if (A_out && h->lazy) { // Julia's f.A on a lazy handle: form the whole matrix once, into a temporary
    double *dA = nullptr, *dc = nullptr;
    CD_TRY(dalloc(&dA, (size_t)h->ld * (size_t)h->p + (size_t)h->p));
    dc = dA + (size_t)h->ld * (size_t)h->p;
    int rcg = launch_gram(h, h->lzX, h->lz_n, (int)h->p, h->lz_ldx, h->lzy, dA, dc, (double)h->lz_n, 1);
    cudaError_t e = cudaStreamSynchronize(h->stream);
    if (rcg == CDGPU_OK && e == cudaSuccess)
      e = cudaMemcpy2D(A_out, h->p * sizeof(double), dA, h->ld * sizeof(double), h->p * sizeof(double), h->p,
                       cudaMemcpyDeviceToHost);
    dfree(dA);
    CD_TRY(rcg);
    CUDA_TRY(e);
  } else if (A_out)
    CUDA_TRY(cudaMemcpy2D(A_out, h->p * sizeof(double), h->dX, h->ld * sizeof(double), h->p * sizeof(double), h->p,
                          cudaMemcpyDeviceToHost));
  if (b_out) CUDA_TRY(cudaMemcpy(b_out, h->dy, h->p * sizeof(double), cudaMemcpyDeviceToHost));
  return CDGPU_OK;
  });
}

// ---------------------------------------------------------------- solves --
static int check_opts(const cdgpu_options *o) {
  if (!o) return cdgpu_set_error(CDGPU_EARG, "null options");
  if (o->maxIter < 0 || o->numSteps < 1) return cdgpu_set_error(CDGPU_EARG, "bad options");
  if (o->randomize < 0 || o->randomize > 1)
    return cdgpu_set_error(CDGPU_EARG, "randomize must be 0 (ordered) or 1 (hash-keyed random permutation)");
  return CDGPU_OK;
}

static int grow(cdgpu_handle_s *h, size_t nlam, size_t outcap, size_t outcols) {
  if ((int64_t)nlam > h->nlam) {
    dfree(h->dlam);
    dfree(h->dstats);
    h->dlam = nullptr;
    h->dstats = nullptr;
    h->nlam = 0;
    CD_TRY(dalloc(&h->dlam, nlam));
    CD_TRY(dalloc(&h->dstats, nlam));
    h->nlam = (int64_t)nlam;
  }
  if ((int64_t)outcap > h->outcap) {
    dfree(h->drowval);
    dfree(h->dnzval);
    h->drowval = nullptr;
    h->dnzval = nullptr;
    h->outcap = 0;
    CD_TRY(dalloc(&h->drowval, outcap));
    CD_TRY(dalloc(&h->dnzval, outcap));
    h->outcap = (int64_t)outcap;
  }
  if ((int64_t)outcols > h->outcols) {
    dfree(h->dcolptr);
    h->dcolptr = nullptr;
    h->outcols = 0;
    CD_TRY(dalloc(&h->dcolptr, outcols + 1));
    h->outcols = (int64_t)outcols;
  }
  return CDGPU_OK;
}

// upload the SparseIterate triple (1-based) as the device list (0-based)
static int upload_iterate(cdgpu_handle_s *h, const double *nzval, const int64_t *nzval2ind, int64_t nnz) {
  if (nnz < 0 || nnz > h->p) return cdgpu_set_error(CDGPU_EDIM, "DimensionMismatch: bad iterate");
  std::vector<int> act((size_t)nnz);
  std::vector<unsigned char> seen((size_t)h->p, 0);
  for (int64_t i = 0; i < nnz; ++i) {
    int64_t k = nzval2ind[i];
    if (k < 1 || k > h->p || seen[(size_t)k - 1]) return cdgpu_set_error(CDGPU_EDIM, "DimensionMismatch: bad iterate");
    seen[(size_t)k - 1] = 1;
    act[(size_t)i] = (int)(k - 1);
  }
  int m = (int)nnz;
  if (nnz) {
    CUDA_TRY(cudaMemcpyAsync(h->dact, act.data(), (size_t)nnz * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->dactval, nzval, (size_t)nnz * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  }
  CUDA_TRY(cudaMemcpyAsync(h->dnact, &m, sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream)); // `act` and `m` are stack/heap temporaries
  return CDGPU_OK;
}
static int download_iterate(cdgpu_handle_s *h, double *nzval, int64_t *nzval2ind, int64_t *nnz) {
  int m = 0;
  CUDA_TRY(cudaMemcpyAsync(&m, h->dnact, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  std::vector<int> act((size_t)m);
  if (m) {
    CUDA_TRY(cudaMemcpyAsync(act.data(), h->dact, (size_t)m * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaMemcpyAsync(nzval, h->dactval, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
  }
  for (int i = 0; i < m; ++i) nzval2ind[i] = (int64_t)act[(size_t)i] + 1;
  *nnz = m;
  return CDGPU_OK;
}
static int set_zero_iterate(cdgpu_handle_s *h) {
  CUDA_TRY(cudaMemsetAsync(h->dnact, 0, sizeof(int), h->stream));
  return CDGPU_OK;
}
static const double *upload_omega(cdgpu_handle_s *h, const double *omega, int *rc) {
  *rc = CDGPU_OK;
  if (!omega) return nullptr;
  cudaError_t e = cudaMemcpyAsync(h->domega, omega, (size_t)h->p * sizeof(double), cudaMemcpyHostToDevice, h->stream);
  if (e != cudaSuccess) *rc = cdgpu_set_error(CDGPU_ECUDA, "H2D omega: %s", cudaGetErrorString(e));
  return h->domega;
}
static void stats_to_abi(const DevStats &d, double ms, cdgpu_stats *s) {
  s->passes = d.passes;
  s->full_passes = d.full_passes;
  s->visits = d.visits;
  s->accepted = d.accepted;
  s->maxH = d.maxH;
  s->converged = d.converged;
  s->outer_iters = d.outer_iters;
  s->sigma = d.sigma;
  s->device_ms = ms;
}
static int flag_status(cdgpu_handle_s *h, int *cols_done) {
  int f[2] = {0, 0};
  CUDA_TRY(cudaMemcpyAsync(f, h->dflag, sizeof f, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (cols_done) *cols_done = f[1];
  if (f[0] == 1) return cdgpu_set_error(CDGPU_ECAP, "path output capacity too small");
  if (f[0] == 2) return cdgpu_set_error(CDGPU_ECAP, "active set larger than the in-CTA engine holds (4096)");
  if (f[0]) return cdgpu_set_error(CDGPU_ECUDA, "kernel reported status %d", f[0]);
  return CDGPU_OK;
}

// log-spaced continuation of coordinate_descent.jl:32-33: numSteps+1 points from lmax to l0
static std::vector<double> continuation(double lmax, double l0, int64_t numSteps) {
  std::vector<double> v;
  double l1 = log(lmax), l2 = log(l0);
  double step = (l2 - l1) / (double)numSteps;
  for (int64_t i = 0; i <= numSteps; ++i) v.push_back(exp(i == numSteps ? l2 : l1 + (double)i * step));
  return v;
}

struct RunCfg {
  const double *lambdas; // host
  int nlambda;
  int accumulate;
  const cdgpu_options *opt;
  const double *domega;
  long long max_hat_s, capacity;
  bool want_path;
  // scaled lasso
  int scaled;
  long long outerMaxIter;
  double outerTol, sigma0;
};

// launches the sweep kernel of the handle's loss over a list of lambdas (device iterate already set)
static int run_sweeps(cdgpu_handle_s *h, const RunCfg &rc) {
  CD_TRY(grow(h, (size_t)rc.nlambda, rc.want_path ? (size_t)rc.capacity : 0, rc.want_path ? (size_t)rc.nlambda : 0));
  CUDA_TRY(cudaMemcpyAsync(h->dlam, rc.lambdas, (size_t)rc.nlambda * sizeof(double), cudaMemcpyHostToDevice,
                           h->stream));
  CUDA_TRY(cudaMemsetAsync(h->dflag, 0, 4 * sizeof(int), h->stream));
  CUDA_TRY(cudaMemsetAsync(h->dflag + 4, 0xff, 8 * sizeof(int), h->stream)); // first-mover words
  CUDA_TRY(cudaMemsetAsync(h->dflag + 12, 0, 4 * sizeof(int), h->stream));
  CUDA_TRY(cudaMemsetAsync(h->dstats, 0, (size_t)rc.nlambda * sizeof(DevStats), h->stream));
  if (rc.want_path) CUDA_TRY(cudaMemsetAsync(h->dcolptr, 0, ((size_t)rc.nlambda + 1) * sizeof(long long), h->stream));
  if (h->kind == CDGPU_LOSS_QUAD) {
    CovArgs a = {};
    a.A = h->dX;
    a.lda = h->ld;
    a.p = (int)h->p;
    a.b = h->dy;
    a.ainv = h->daux;
    a.omega = rc.domega;
    a.Ax = h->dstate;
    a.beta = h->dbeta;
    a.act = h->dact;
    a.actval = h->dactval;
    a.nact = h->dnact;
    a.inlist = h->dinlist;
    a.scr = h->dscr;
    a.iscr = h->discr;
    a.bscr = h->dbscr;
    a.chain_scr = h->dchain;
    a.lambdas = h->dlam;
    a.nlambda = rc.nlambda;
    a.accumulate = rc.accumulate;
    a.maxIter = rc.opt->maxIter;
    a.optTol = rc.opt->optTol;
    a.randomize = rc.opt->randomize;
    a.seed = rc.opt->seed;
    a.max_hat_s = rc.max_hat_s;
    a.colptr = rc.want_path ? h->dcolptr : nullptr;
    a.rowval = h->drowval;
    a.nzval = h->dnzval;
    a.capacity = rc.capacity;
    a.flag = h->dflag;
    a.stats = h->dstats;
    const bool prof = getenv("CDGPU_PROFILE") != nullptr;
    a.prof = prof ? reinterpret_cast<long long *>(h->dscr + 11 * (size_t)h->p) : nullptr;
    a.events_only = getenv("CDGPU_COV_EVENTS") != nullptr && atoi(getenv("CDGPU_COV_EVENTS")) != 0;
    if (!h->lazy && getenv("CDGPU_IDENTITY_SLOT")) { // diagnostics: an eager handle through the column-slot indirection
      std::vector<int> id((size_t)h->p);
      for (int64_t j = 0; j < h->p; ++j) id[(size_t)j] = (int)j;
      int *dslot_id = h->discr + 9 * (size_t)h->p; // [9p, 10p): free between launches
      CUDA_TRY(cudaMemcpyAsync(dslot_id, id.data(), (size_t)h->p * sizeof(int), cudaMemcpyHostToDevice, h->stream));
      CUDA_TRY(cudaStreamSynchronize(h->stream));
      a.colslot = dslot_id;
    }
    if (h->lazy) {
      // columns of the members handed in (warm start) must exist; a fresh handle also forms its first batch now: the
      // coordinates with the largest |b_j|/omega_j are the ones that enter first along a path
      a.colslot = h->dslot;
      a.resume = h->dresume;
      CD_TRY(lazy_spec_commit(h)); // left in flight by the previous solve
      h->lz_batches = h->lz_pauses = 0;
      h->lz_form_ms = 0.0;
      h->lz_form_cols = 0;
      int m0 = 0;
      CUDA_TRY(cudaMemcpyAsync(&m0, h->dnact, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
      CUDA_TRY(cudaStreamSynchronize(h->stream));
      std::vector<int> need((size_t)m0);
      if (m0) {
        CUDA_TRY(cudaMemcpyAsync(need.data(), h->dact, (size_t)m0 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
      }
      CD_TRY(lazy_ensure(h, need, rc.domega, h->lz_used == 0 && !getenv("CDGPU_LAZY_NO_PREFETCH")));
      a.A = h->dX; // the cache may have moved (lazy_reserve)
      CUDA_TRY(cudaMemsetAsync(h->dresume, 0, sizeof(CovResume), h->stream));
    }
    if (prof) CUDA_TRY(cudaMemsetAsync(a.prof, 0, 24 * sizeof(long long), h->stream));
    h->sweep_ms = 0.0;
    CD_TRY(launch_cov_init(h, a.A, a.lda, a.p, a.act, a.actval, a.nact, a.Ax, a.beta, a.inlist, a.colslot));
    CUDA_TRY(cudaEventRecord(h->sw_ev0, h->stream));
    CD_TRY(launch_cov_path(h, a));
    CUDA_TRY(cudaEventRecord(h->sw_ev1, h->stream));
    if (h->lazy) CD_TRY(lazy_spec_start(h));
    while (h->lazy) { // paused at an entering coordinate without a column: form a batch, resume
      int f[2] = {0, 0};
      CUDA_TRY(cudaMemcpyAsync(f, h->dflag, sizeof f, cudaMemcpyDeviceToHost, h->stream));
      CUDA_TRY(cudaStreamSynchronize(h->stream));
      {
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, h->sw_ev0, h->sw_ev1));
        h->sweep_ms += ms;
      }
      if (f[0] != 3) break;
      CovResume R;
      CUDA_TRY(cudaMemcpy(&R, h->dresume, sizeof R, cudaMemcpyDeviceToHost));
      if (R.need_k < 0 || R.need_k >= h->p) return cdgpu_set_error(CDGPU_ECUDA, "lazy covariance: bad column request %d", R.need_k);
      h->lz_pauses += 1;
      CD_TRY(lazy_spec_commit(h)); // the batch formed in the background while the kernel ran, if any
      std::vector<int> need(1, R.need_k);
      const bool missed = h->hslot[R.need_k] < 0; // speculation did not cover it: form a batch now, speculate again afterwards
      CD_TRY(lazy_ensure(h, need, rc.domega, !getenv("CDGPU_LAZY_NO_PREFETCH")));
      a.A = h->dX; // the cache may have moved (lazy_reserve)
      const int one = 1;
      CUDA_TRY(cudaMemcpyAsync(&h->dresume->valid, &one, sizeof(int), cudaMemcpyHostToDevice, h->stream));
      CUDA_TRY(cudaStreamSynchronize(h->stream)); // `one` is a stack temporary
      CUDA_TRY(cudaEventRecord(h->sw_ev0, h->stream));
      CD_TRY(launch_cov_path(h, a));
      CUDA_TRY(cudaEventRecord(h->sw_ev1, h->stream));
      if (missed) CD_TRY(lazy_spec_start(h));
    }
    // a speculative batch still in flight is committed by the next solve on this handle (or dropped with the handle)
    if (!h->lazy) h->sweep_pending = true; // eager: one launch, its time is read when somebody asks (no extra sync here)
    if (prof) {
      long long pf[24];
      CUDA_TRY(cudaMemcpyAsync(pf, a.prof, sizeof pf, cudaMemcpyDeviceToHost, h->stream));
      CUDA_TRY(cudaStreamSynchronize(h->stream));
      fprintf(stderr,
              "[cdgpu profile] cov path: total %.3f Mcyc | full passes %.3f (events %lld) | list update %.3f | "
              "active engine %.3f (steps %lld) | refresh %.3f || full-pass rounds: scan+publish %.3f, cluster.sync %.3f, "
              "apply %.3f\n",
              pf[6] * 1e-6, pf[0] * 1e-6, pf[4], pf[1] * 1e-6, pf[2] * 1e-6, pf[5], pf[3] * 1e-6, pf[7] * 1e-6, pf[8] * 1e-6,
              pf[9] * 1e-6);
      fprintf(stderr,
              "[cdgpu profile]   chain engine: warp0 panel %.3f chain %.3f barrier %.3f pass-ends %.3f | workers stage %.3f apply "
              "%.3f barrier %.3f Mcyc\n",
              pf[16] * 1e-6, pf[17] * 1e-6, pf[18] * 1e-6, pf[19] * 1e-6, pf[20] * 1e-6, pf[21] * 1e-6, pf[22] * 1e-6);
      fprintf(stderr, "[cdgpu profile]   verify sweep (thread 0 of CTA 0): set-up %.3f loop %.3f tests %.3f tail %.3f Mcyc "
                      "| screened: CTA 0 promoted %lld rows, CURRENT rows per round %.1f over %lld rounds, resyncs %lld\n",
              pf[10] * 1e-6, pf[11] * 1e-6, pf[12] * 1e-6, pf[13] * 1e-6, pf[10], pf[12] ? (double)pf[11] / (double)pf[12] : 0.0, pf[12],
              pf[13]);
      fprintf(stderr, "[cdgpu profile]   screened round on CTA 0: copy/sync %.3f, bound tests %.3f, catch-up %.3f Mcyc\n", pf[14] * 1e-6,
              pf[15] * 1e-6, pf[23] * 1e-6);
    }
  } else {
    NaiveArgs a = {};
    a.kind = h->kind;
    a.X = h->dX;
    a.ldx = h->ld;
    a.n = (int)h->n;
    a.p = (int)h->p;
    a.y = h->dy;
    a.w = h->dw;
    a.colsq = h->daux;
    a.omega = rc.domega;
    a.r = h->dstate;
    a.beta = h->dbeta;
    a.act = h->dact;
    a.actval = h->dactval;
    a.nact = h->dnact;
    a.inlist = h->dinlist;
    a.scr = h->dscr;
    a.iscr = h->discr;
    a.bscr = h->dbscr;
    a.chain_scr = h->dchain;
    a.lambdas = h->dlam;
    a.nlambda = rc.nlambda;
    a.accumulate = rc.accumulate;
    a.maxIter = rc.opt->maxIter;
    a.optTol = rc.opt->optTol;
    a.randomize = rc.opt->randomize;
    a.seed = rc.opt->seed;
    a.max_hat_s = rc.max_hat_s;
    a.colptr = rc.want_path ? h->dcolptr : nullptr;
    a.rowval = h->drowval;
    a.nzval = h->dnzval;
    a.capacity = rc.capacity;
    a.flag = h->dflag;
    a.stats = h->dstats;
    const size_t gcap = cd_gram_cap((size_t)h->p);
    if (!h->dgram && !getenv("CDGPU_NAIVE_NO_GRAM_ENGINE")) CD_TRY(dalloc(&h->dgram, gcap * gcap + gcap));
    a.gram = getenv("CDGPU_NAIVE_NO_GRAM_ENGINE") ? nullptr : h->dgram;
    a.gram_cap = (int)gcap;
    if (const char *env = getenv("CDGPU_NAIVE_GCAP")) a.gram_cap = std::max(256, std::min((int)gcap, atoi(env))); // diagnostics
    a.multi_ok = 192; // smallest active set handed to the 16-CTA team engine (C1 lambda=0.05, 230 entries: 5.9 -> 5.45 ms against 384)
    if (const char *env = getenv("CDGPU_MULTI_MIN")) a.multi_ok = std::max(64, atoi(env));
    if (const char *env = getenv("CDGPU_NAIVE_MULTI")) a.multi_ok = atoi(env) != 0 ? a.multi_ok : 0;
    a.pipeline = 1;
    if (const char *env = getenv("CDGPU_NAIVE_PIPELINE")) a.pipeline = atoi(env) != 0;
    a.plan = 1;
    if (const char *env = getenv("CDGPU_NAIVE_PLAN")) a.plan = atoi(env) != 0;
    a.rsnap = h->drsnap;
    a.dense = 1;
    if (const char *env = getenv("CDGPU_NAIVE_DENSE")) a.dense = atoi(env) != 0;
    a.replan = 0; // only in builds with -DCDGPU_WITH_REPLAN (naive_sweep.cu: full_pass)
    if (const char *env = getenv("CDGPU_NAIVE_REPLAN")) a.replan = atoi(env) != 0;
    a.scaled = rc.scaled;
    a.outerMaxIter = rc.outerMaxIter;
    a.outerTol = rc.outerTol;
    a.sigma0 = rc.sigma0;
    const bool prof = getenv("CDGPU_PROFILE") != nullptr;
    a.prof = prof ? reinterpret_cast<long long *>(h->dscr + cd_scr_tail((size_t)h->p, (size_t)h->n) + 16 * (size_t)h->p + 8) : nullptr; // 32 doubles behind the round buffers
    CD_TRY(launch_naive_init(h, a));
    CD_TRY(launch_naive_path(h, a));
    if (prof) {
      long long pf[32];
      CUDA_TRY(cudaMemcpyAsync(pf, a.prof, sizeof pf, cudaMemcpyDeviceToHost, h->stream));
      CUDA_TRY(cudaStreamSynchronize(h->stream));
      fprintf(stderr,
              "[cdgpu profile] naive path (CTA 0): total %.3f Mcyc | full-pass rounds %lld: column dots %.3f, grid.sync "
              "%.3f, scan %.3f, apply %.3f | list update %.3f | active phase %.3f (forming the active Gram %.3f) | member plan %.3f\n",
              pf[6] * 1e-6, pf[7], pf[0] * 1e-6, pf[1] * 1e-6, pf[2] * 1e-6, pf[3] * 1e-6, pf[4] * 1e-6, pf[5] * 1e-6, pf[9] * 1e-6,
              pf[8] * 1e-6);
      fprintf(stderr,
              "[cdgpu profile]   chain engine: warp0 panel %.3f chain %.3f barrier %.3f pass-ends %.3f | workers stage %.3f apply (team: "
              "warp 0 waiting for owners) %.3f barrier %.3f | team engine (counters with -DCDGPU_CHAIN_PROF): apply on CTA 1 %.3f; whole call %.3f, of which publish %.3f Mcyc\n",
              pf[16] * 1e-6, pf[17] * 1e-6, pf[18] * 1e-6, pf[19] * 1e-6, pf[20] * 1e-6, pf[21] * 1e-6, pf[22] * 1e-6, pf[23] * 1e-6,
              pf[10] * 1e-6, pf[11] * 1e-6);
      fprintf(stderr,
              "[cdgpu profile]   owner warp (CTA 1, thread 0; -DCDGPU_CHAIN_PROF): wait g %.3f, wait h %.3f, 32 updates %.3f, tagged store %.3f, "
              "prefetch %.3f Mcyc\n",
              pf[24] * 1e-6, pf[25] * 1e-6, pf[26] * 1e-6, pf[27] * 1e-6, pf[28] * 1e-6);
    }
  }
  return CDGPU_OK;
}

static int lambda_max_dev(cdgpu_handle_s *h, const double *domega, double *out_host) {
  double *dout = h->dscr; // first slot of the scratch
  if (h->kind == CDGPU_LOSS_QUAD) {
    CD_TRY(launch_lambda_max_quad(h, h->dy, domega, (int)h->p, dout));
  } else {
    CD_TRY(launch_lambda_max_naive(h, h->kind, h->dX, h->ld, (int)h->n, (int)h->p, h->dy, h->dw, domega, h->dscr + 8,
                                   dout));
  }
  CUDA_TRY(cudaMemcpyAsync(out_host, dout, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return CDGPU_OK;
}

API int cdgpu_solve(cdgpu_handle h, double lambda0, const double *omega, const cdgpu_options *opt, double *nzval,
                    int64_t *nzval2ind, int64_t *nnz, cdgpu_stats *stats) {
  return api_guard([&]() -> int {
  if (!h || !nzval || !nzval2ind || !nnz) return cdgpu_set_error(CDGPU_EARG, "null pointer");
  CD_TRY(check_opts(opt));
  CUDA_TRY(cudaSetDevice(h->device));
  if (h->tall) { // tall LS / WLS: the inner covariance handle solves, then f.r is formed
    CD_TRY(cdgpu_solve(h->tall, lambda0, omega, opt, nzval, nzval2ind, nnz, stats));
    return tall_residual(h);
  }
  int rc;
  const double *domega = upload_omega(h, omega, &rc);
  CD_TRY(rc);
  CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
  RunCfg cfg = {};
  cfg.opt = opt;
  cfg.domega = domega;
  cfg.accumulate = 1;
  cfg.max_hat_s = -1;
  std::vector<double> lams;
  if (opt->warmStart) {
    CD_TRY(upload_iterate(h, nzval, nzval2ind, *nnz));
    lams.push_back(lambda0);
  } else {
    CD_TRY(set_zero_iterate(h)); // fill!(x, 0)  coordinate_descent.jl:25
    double lmax = 0.0;
    CD_TRY(lambda_max_dev(h, domega, &lmax));
    lams = continuation(lmax, lambda0, opt->numSteps);
  }
  cfg.lambdas = lams.data();
  cfg.nlambda = (int)lams.size();
  CD_TRY(run_sweeps(h, cfg));
  CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
  CD_TRY(flag_status(h, nullptr));
  CD_TRY(download_iterate(h, nzval, nzval2ind, nnz));
  if (stats) {
    DevStats d;
    CUDA_TRY(cudaMemcpy(&d, h->dstats, sizeof d, cudaMemcpyDeviceToHost));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    stats_to_abi(d, ms, stats);
  }
  return CDGPU_OK;
  });
}

API int cdgpu_path(cdgpu_handle h, const double *lambda, int64_t m, const double *omega, const cdgpu_options *opt,
                   int64_t max_hat_s, int64_t capacity, int64_t *colptr, int64_t *rowval, double *nzval,
                   int64_t *m_done, cdgpu_stats *stats) {
  return api_guard([&]() -> int {
  if (!h || !lambda || !colptr || !rowval || !nzval || !m_done) return cdgpu_set_error(CDGPU_EARG, "null pointer");
  if (m < 0 || m > 0x7fffffff || capacity < 0) return cdgpu_set_error(CDGPU_EARG, "bad path length or capacity");
  CD_TRY(check_opts(opt));
  CUDA_TRY(cudaSetDevice(h->device));
  colptr[0] = 0;
  *m_done = 0;
  if (m == 0) return CDGPU_OK;
  if (h->tall) {
    CD_TRY(cdgpu_path(h->tall, lambda, m, omega, opt, max_hat_s, capacity, colptr, rowval, nzval, m_done, stats));
    return tall_residual(h);
  }
  if (!opt->warmStart) {
    // LassoPath with warmStart=false re-runs the internal continuation for every lambda
    // (coordinate_descent.jl:23-37): run it point by point.
    int64_t off = 0;
    std::vector<double> nz((size_t)h->p);
    std::vector<int64_t> ind((size_t)h->p);
    for (int64_t i = 0; i < m; ++i) {
      int64_t nn = 0;
      CD_TRY(cdgpu_solve(h, lambda[i], omega, opt, nz.data(), ind.data(), &nn, stats ? stats + i : nullptr));
      if (off + nn > capacity) return cdgpu_set_error(CDGPU_ECAP, "path output capacity too small");
      memcpy(rowval + off, ind.data(), (size_t)nn * sizeof(int64_t));
      memcpy(nzval + off, nz.data(), (size_t)nn * sizeof(double));
      off += nn;
      colptr[i + 1] = off;
      *m_done = i + 1;
      if (max_hat_s >= 0 && nn > max_hat_s) break;
    }
    return CDGPU_OK;
  }
  int rc;
  const double *domega = upload_omega(h, omega, &rc);
  CD_TRY(rc);
  CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
  CD_TRY(set_zero_iterate(h)); // x = SparseIterate(T, p)  lasso.jl:244
  RunCfg cfg = {};
  cfg.opt = opt;
  cfg.domega = domega;
  cfg.accumulate = 0;
  cfg.max_hat_s = max_hat_s;
  cfg.capacity = capacity;
  cfg.want_path = true;
  cfg.lambdas = lambda;
  cfg.nlambda = (int)m;
  CD_TRY(run_sweeps(h, cfg));
  CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
  int cols = 0;
  int st = flag_status(h, &cols);
  float ms = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  if (cols > 0) {
    static_assert(sizeof(long long) == sizeof(int64_t), "int64");
    CUDA_TRY(cudaMemcpy(colptr, h->dcolptr, ((size_t)cols + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost));
    colptr[0] = 0;
    const int64_t tot = colptr[cols];
    if (tot > 0) {
      CUDA_TRY(cudaMemcpy(rowval, h->drowval, (size_t)tot * sizeof(int64_t), cudaMemcpyDeviceToHost));
      CUDA_TRY(cudaMemcpy(nzval, h->dnzval, (size_t)tot * sizeof(double), cudaMemcpyDeviceToHost));
    }
    if (stats) {
      std::vector<DevStats> d((size_t)cols);
      CUDA_TRY(cudaMemcpy(d.data(), h->dstats, (size_t)cols * sizeof(DevStats), cudaMemcpyDeviceToHost));
      for (int i = 0; i < cols; ++i) stats_to_abi(d[(size_t)i], i == 0 ? (double)ms : 0.0, stats + i);
    }
  }
  *m_done = cols;
  return st;
  });
}

// _findInitSigma! (utils.jl:60-77): the s columns most correlated with y (device: |X'y|), a tiny
// least-squares fit of y on them (host: Householder QR on n x s, as LAPACK's `\` does) and the
// corrected standard deviation of its residuals.  One host round trip at initialisation only.
static int screening_sigma(cdgpu_handle_s *h, int64_t sinit, double *sigma) {
  const int64_t n = h->n, p = h->p;
  int64_t s = sinit < 1 ? 1 : (sinit > p ? p : sinit);
  double *dcorr = h->dscr + 8;
  CD_TRY(launch_abs_xty(h, h->dX, h->ld, (int)n, (int)p, h->dy, dcorr));
  std::vector<double> corr((size_t)p), y((size_t)n);
  CUDA_TRY(cudaMemcpyAsync(corr.data(), dcorr, (size_t)p * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaMemcpyAsync(y.data(), h->dy, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  std::vector<double> sorted(corr);
  std::nth_element(sorted.begin(), sorted.begin() + (s - 1), sorted.end(), [](double a, double b) { return a > b; });
  const double thr = sorted[(size_t)s - 1]; // nlargest(s, storage)[end]
  std::vector<int64_t> S;
  for (int64_t j = 0; j < p; ++j)
    if (corr[(size_t)j] >= thr) S.push_back(j); // storage .>= thr (ties may select more than s)
  const int64_t cnt = (int64_t)S.size();
  std::vector<double> Xs((size_t)(n * cnt)), Q, rhs(y), beta((size_t)cnt, 0.0);
  for (int64_t q = 0; q < cnt; ++q)
    CUDA_TRY(cudaMemcpyAsync(Xs.data() + q * n, h->dX + S[(size_t)q] * h->ld, (size_t)n * sizeof(double),
                             cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  Q = Xs;
  for (int64_t j = 0; j < cnt; ++j) { // Householder QR applied to [Q | rhs]
    double *c = Q.data() + j * n;
    double nrm = 0.0;
    for (int64_t i = j; i < n; ++i) nrm += c[i] * c[i];
    nrm = sqrt(nrm);
    if (nrm == 0.0) continue;
    const double alpha = c[j] > 0 ? -nrm : nrm;
    const double v0 = c[j] - alpha;
    double vn2 = v0 * v0;
    for (int64_t i = j + 1; i < n; ++i) vn2 += c[i] * c[i];
    c[j] = v0;
    for (int64_t jj = j + 1; jj <= cnt; ++jj) {
      double *t = jj < cnt ? Q.data() + jj * n : rhs.data();
      double d = 0.0;
      for (int64_t i = j; i < n; ++i) d += c[i] * t[i];
      d = 2.0 * d / vn2;
      for (int64_t i = j; i < n; ++i) t[i] -= d * c[i];
    }
    c[j] = alpha;
  }
  for (int64_t j = cnt - 1; j >= 0; --j) {
    double v = rhs[(size_t)j];
    for (int64_t jj = j + 1; jj < cnt; ++jj) v -= Q[(size_t)(j + jj * n)] * beta[(size_t)jj];
    beta[(size_t)j] = v / Q[(size_t)(j + j * n)];
  }
  double mean = 0.0;
  std::vector<double> r((size_t)n);
  for (int64_t i = 0; i < n; ++i) {
    double v = 0.0;
    for (int64_t q = 0; q < cnt; ++q) v += Xs[(size_t)(i + q * n)] * beta[(size_t)q];
    r[(size_t)i] = y[(size_t)i] - v;
    mean += r[(size_t)i];
  }
  mean /= (double)n;
  double var = 0.0;
  for (int64_t i = 0; i < n; ++i) var += (r[(size_t)i] - mean) * (r[(size_t)i] - mean);
  *sigma = sqrt(var / (double)(n - 1)); // Statistics.std
  return CDGPU_OK;
}

// scaledLasso! with optionsCD.warmStart == false: every outer iteration is a cold coordinateDescent! (fill!(x,0) +
// the numSteps continuation from lambda_max, coordinate_descent.jl:23-37 called from lasso.jl:132-133), so the sigma
// loop cannot stay inside one launch: one launch per outer iteration, sigma from the device residual in between.
static double host_std(const std::vector<double> &r) { // Statistics.std (corrected)
  const size_t n = r.size();
  double mean = 0.0, var = 0.0;
  for (double v : r) mean += v;
  mean /= (double)n;
  for (double v : r) var += (v - mean) * (v - mean);
  return sqrt(var / (double)(n - 1));
}
static int scaled_solve_cold(cdgpu_handle_s *h, double lambda, const double *domega, const cdgpu_iter_options *opt,
                             double sigma0, double *nzval, int64_t *nzval2ind, int64_t *nnz, double *sigma_out,
                             cdgpu_stats *stats) {
  const size_t n = (size_t)h->n;
  std::vector<double> r(n);
  auto fetch_r = [&]() -> int {
    CUDA_TRY(cudaMemcpyAsync(r.data(), h->dstate, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return CDGPU_OK;
  };
  RunCfg cfg = {};
  cfg.domega = domega;
  cfg.accumulate = 1;
  cfg.max_hat_s = -1;
  double sigma = sigma0;
  if (opt->initProcedure == CDGPU_INIT_WARMSTART) { // initialize!(f, x); sigma = std(f.r)   lasso.jl:124-126
    cdgpu_options o0 = opt->optionsCD;
    o0.maxIter = 0;
    cfg.opt = &o0;
    cfg.lambdas = &lambda;
    cfg.nlambda = 1;
    CD_TRY(run_sweeps(h, cfg));
    CD_TRY(fetch_r());
    sigma = host_std(r);
  }
  cfg.opt = &opt->optionsCD;
  DevStats tot = {};
  int outer = 0;
  for (int64_t iter = 1; iter <= opt->maxIter; ++iter) {
    outer = (int)iter;
    CD_TRY(set_zero_iterate(h));
    double lmax = 0.0;
    CD_TRY(lambda_max_dev(h, domega, &lmax));
    const std::vector<double> lams = continuation(lmax, lambda * sigma, opt->optionsCD.numSteps);
    cfg.lambdas = lams.data();
    cfg.nlambda = (int)lams.size();
    CD_TRY(run_sweeps(h, cfg));
    CD_TRY(flag_status(h, nullptr));
    DevStats d;
    CUDA_TRY(cudaMemcpy(&d, h->dstats, sizeof d, cudaMemcpyDeviceToHost));
    tot.passes += d.passes;
    tot.full_passes += d.full_passes;
    tot.visits += d.visits;
    tot.accepted += d.accepted;
    tot.maxH = d.maxH;
    tot.converged = d.converged;
    CD_TRY(fetch_r());
    double ss = 0.0;
    for (double v : r) ss += v * v;
    const double snew = sqrt(ss / (double)n); // lasso.jl:134
    if (fabs(snew - sigma) / sigma < opt->optTol) break;
    sigma = snew;
  }
  CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
  CD_TRY(download_iterate(h, nzval, nzval2ind, nnz));
  tot.outer_iters = outer;
  tot.sigma = sigma;
  float ms = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  if (stats) stats_to_abi(tot, ms, stats);
  if (sigma_out) *sigma_out = n > 1 ? host_std(r) : 0.0; // std(f.r)  lasso.jl:143
  return CDGPU_OK;
}

// scaledLasso! (lasso.jl:107-144) on a tall LS handle: the inner covariance handle runs each coordinateDescent!, sigma is
// re-estimated from the residual formed on the device (one launch per outer iteration instead of one for the loop)
static int scaled_solve_tall(cdgpu_handle_s *h, double lambda, const double *omega, const cdgpu_iter_options *opt,
                             double sigma0, double *nzval, int64_t *nzval2ind, int64_t *nnz, double *sigma_out,
                             cdgpu_stats *stats) {
  const size_t n = (size_t)h->n;
  std::vector<double> r(n);
  auto fetch_r = [&]() -> int {
    CD_TRY(tall_residual(h));
    CUDA_TRY(cudaMemcpy(r.data(), h->dstate, n * sizeof(double), cudaMemcpyDeviceToHost));
    return CDGPU_OK;
  };
  const auto t0 = std::chrono::steady_clock::now();
  double sigma = sigma0;
  if (opt->initProcedure == CDGPU_INIT_WARMSTART) { // initialize!(f, x); sigma = std(f.r)   lasso.jl:124-126
    CD_TRY(upload_iterate(h->tall, nzval, nzval2ind, *nnz));
    CD_TRY(fetch_r());
    sigma = host_std(r);
  }
  cdgpu_stats tot = {}, st = {};
  int outer = 0;
  for (int64_t iter = 1; iter <= opt->maxIter; ++iter) {
    outer = (int)iter;
    CD_TRY(cdgpu_solve(h->tall, lambda * sigma, omega, &opt->optionsCD, nzval, nzval2ind, nnz, &st));
    tot.passes += st.passes;
    tot.full_passes += st.full_passes;
    tot.visits += st.visits;
    tot.accepted += st.accepted;
    tot.maxH = st.maxH;
    tot.converged = st.converged;
    tot.device_ms += st.device_ms;
    CD_TRY(fetch_r());
    double ss = 0.0;
    for (double v : r) ss += v * v;
    const double snew = sqrt(ss / (double)n); // lasso.jl:134
    if (fabs(snew - sigma) / sigma < opt->optTol) break;
    sigma = snew;
  }
  tot.outer_iters = outer;
  tot.sigma = sigma;
  tot.device_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  if (stats) *stats = tot;
  if (sigma_out) *sigma_out = n > 1 ? host_std(r) : 0.0; // std(f.r)  lasso.jl:143
  return CDGPU_OK;
}

API int cdgpu_scaled_solve(cdgpu_handle h, double lambda, const double *omega, const cdgpu_iter_options *opt,
                           double *nzval, int64_t *nzval2ind, int64_t *nnz, double *sigma_out, cdgpu_stats *stats) {
  return api_guard([&]() -> int {
  if (!h || !opt || !omega || !nzval || !nzval2ind || !nnz) return cdgpu_set_error(CDGPU_EARG, "null pointer");
  if (h->kind != CDGPU_LOSS_LS)
    return cdgpu_set_error(CDGPU_EARG, "scaledLasso! needs a CDLeastSquaresLoss handle (lasso.jl:117)");
  CD_TRY(check_opts(&opt->optionsCD));
  if (opt->initProcedure != CDGPU_INIT_STD && opt->initProcedure != CDGPU_INIT_WARMSTART &&
      opt->initProcedure != CDGPU_INIT_SCREENING)
    return cdgpu_set_error(CDGPU_EARG, "ArgumentError: Incorrect initialization Symbol");
  CUDA_TRY(cudaSetDevice(h->device));
  double sigma0 = opt->sigma_init;
  if (opt->initProcedure == CDGPU_INIT_SCREENING) CD_TRY(screening_sigma(h, opt->sinit, &sigma0));
  if (h->tall) return scaled_solve_tall(h, lambda, omega, opt, sigma0, nzval, nzval2ind, nnz, sigma_out, stats);
  int rc;
  const double *domega = upload_omega(h, omega, &rc);
  CD_TRY(rc);
  CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
  CD_TRY(upload_iterate(h, nzval, nzval2ind, *nnz));
  if (!opt->optionsCD.warmStart)
    return scaled_solve_cold(h, lambda, domega, opt, sigma0, nzval, nzval2ind, nnz, sigma_out, stats);
  RunCfg cfg = {};
  cfg.opt = &opt->optionsCD;
  cfg.domega = domega;
  cfg.accumulate = 1;
  cfg.max_hat_s = -1;
  cfg.lambdas = &lambda;
  cfg.nlambda = 1;
  cfg.scaled = opt->initProcedure == CDGPU_INIT_WARMSTART ? 2 : 1; // 2: sigma0 = std(r) computed on the device
  cfg.outerMaxIter = opt->maxIter;
  cfg.outerTol = opt->optTol;
  cfg.sigma0 = sigma0;
  CD_TRY(run_sweeps(h, cfg));
  CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
  CD_TRY(flag_status(h, nullptr));
  CD_TRY(download_iterate(h, nzval, nzval2ind, nnz));
  DevStats d;
  CUDA_TRY(cudaMemcpy(&d, h->dstats, sizeof d, cudaMemcpyDeviceToHost));
  float ms = 0.f;
  CUDA_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  if (stats) stats_to_abi(d, ms, stats);
  if (sigma_out) { // std(f.r) of lasso.jl:143, written by the kernel next to the flags
    double s[2];
    CUDA_TRY(cudaMemcpy(s, h->dscr, sizeof s, cudaMemcpyDeviceToHost));
    *sigma_out = s[0];
  }
  return CDGPU_OK;
  });
}

API int cdgpu_state(cdgpu_handle h, double *out) {
  return api_guard([&]() -> int {
  if (!h || !out) return cdgpu_set_error(CDGPU_EARG, "null pointer");
  CUDA_TRY(cudaSetDevice(h->device));
  const size_t cnt = (size_t)(h->kind == CDGPU_LOSS_QUAD ? h->p : h->n);
  CUDA_TRY(cudaMemcpy(out, h->dstate, cnt * sizeof(double), cudaMemcpyDeviceToHost));
  return CDGPU_OK;
  });
}

API int cdgpu_stdx(cdgpu_handle h, const double *w, double *out) {
  return api_guard([&]() -> int {
  if (!h || !out) return cdgpu_set_error(CDGPU_EARG, "null pointer");
  CUDA_TRY(cudaSetDevice(h->device));
  if (h->kind == CDGPU_LOSS_QUAD) { // A = X'X/n  =>  _stdX!(X)_j = sqrt(A_jj)
    if (w) return cdgpu_set_error(CDGPU_EARG, "weighted _stdX! needs a naive-form handle");
    double *dout = h->dscr + 8;
    if (h->lazy)
      CD_TRY(launch_sqrt_vec(h, h->ddiag, (int)h->p, dout));
    else
      CD_TRY(launch_diag_sqrt(h, h->dX, h->ld, (int)h->p, dout));
    CUDA_TRY(cudaMemcpyAsync(out, dout, (size_t)h->p * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return CDGPU_OK;
  }
  double *dw = nullptr;
  if (w) {
    dw = h->dscr + 8;
    CUDA_TRY(cudaMemcpyAsync(dw, w, (size_t)h->n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  }
  double *dout = h->dscr + 8 + h->n;
  CD_TRY(launch_colsq(h, h->dX, h->ld, (int)h->n, (int)h->p, dw, dout, true));
  CUDA_TRY(cudaMemcpyAsync(out, dout, (size_t)h->p * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return CDGPU_OK;
  });
}

// refitLassoPath (lasso.jl:208-225): coefficients of the least-squares fit on the columns `support` (1-based, ns of
// them) of the handle's design: X[:, S] \ y for a naive-form handle ([W]-weighted for CDWeightedLSLoss), the
// equivalent A[S,S] \ (-b[S]) for a covariance-form handle.  Normal equations + Cholesky on the device.
API int cdgpu_refit(cdgpu_handle h, const int64_t *support, int64_t ns, double *coef_out) {
  return api_guard([&]() -> int {
  if (!h || (ns > 0 && (!support || !coef_out))) return cdgpu_set_error(CDGPU_EARG, "null pointer");
  if (ns < 0 || ns > h->p) return cdgpu_set_error(CDGPU_EDIM, "DimensionMismatch");
  if (ns == 0) return CDGPU_OK;
  if (ns > (int64_t)cd_gram_cap((size_t)h->p))
    return cdgpu_set_error(CDGPU_ECAP, "refit handles supports of at most %d columns", (int)cd_gram_cap((size_t)h->p));
  CUDA_TRY(cudaSetDevice(h->device));
  std::vector<int> s0((size_t)ns);
  for (int64_t i = 0; i < ns; ++i) {
    if (support[i] < 1 || support[i] > h->p) return cdgpu_set_error(CDGPU_EDIM, "BoundsError: support index out of range");
    s0[(size_t)i] = (int)(support[i] - 1);
  }
  if (!h->dgram) {
    const size_t gcap = cd_gram_cap((size_t)h->p);
    CD_TRY(dalloc(&h->dgram, gcap * gcap + gcap));
  }
  int *dS = h->discr; // 8p ints of scratch
  CUDA_TRY(cudaMemcpyAsync(dS, s0.data(), (size_t)ns * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream)); // s0 is a host temporary
  if (h->lazy) { // covariance columns may be missing: the same normal equations straight from the resident data
    cdgpu_handle_s t = *h;
    t.kind = CDGPU_LOSS_LS;
    t.dX = const_cast<double *>(h->lzX);
    t.ld = h->lz_ldx;
    t.n = h->lz_n;
    t.dy = const_cast<double *>(h->lzy);
    CD_TRY(launch_refit(&t, dS, (int)ns, h->dgram, h->dflag));
  } else {
    CD_TRY(launch_refit(h, dS, (int)ns, h->dgram, h->dflag));
  }
  const int ld = ((int)ns + 1) & ~1;
  int flag = 0;
  CUDA_TRY(cudaMemcpyAsync(coef_out, h->dgram + (size_t)ld * ns, (size_t)ns * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaMemcpyAsync(&flag, h->dflag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  CUDA_TRY(cudaMemsetAsync(h->dflag, 0, sizeof(int), h->stream));
  if (flag) return cdgpu_set_error(CDGPU_EARG, "SingularException: the selected columns are not linearly independent");
  return CDGPU_OK;
  });
}

API int cdgpu_lambda_max(cdgpu_handle h, const double *omega, double *out) {
  return api_guard([&]() -> int {
  if (!h || !out) return cdgpu_set_error(CDGPU_EARG, "null pointer");
  CUDA_TRY(cudaSetDevice(h->device));
  int rc;
  const double *domega = upload_omega(h, omega, &rc);
  CD_TRY(rc);
  return lambda_max_dev(h, domega, out);
  });
}
