// common.cuh — shared declarations of libcdgpu (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cdgpu.h"

// ---------------------------------------------------------------- errors --
int cdgpu_set_error(int code, const char *fmt, ...);
#define CUDA_TRY(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess)                                                                               \
      return cdgpu_set_error(_e == cudaErrorMemoryAllocation ? CDGPU_ENOMEM : CDGPU_ECUDA, "%s: %s (%s:%d)", #expr, \
                             cudaGetErrorString(_e), __FILE__, __LINE__);                                \
  } while (0)
#define CD_TRY(expr)       \
  do {                     \
    int _rc = (expr);      \
    if (_rc) return _rc;   \
  } while (0)

// No C++ exception crosses the C ABI (include/cdgpu.h): every entry point that can allocate runs inside this guard.
#ifdef __cplusplus
#include <new>
template <class F>
static inline int api_guard(F &&f) noexcept {
  try {
    return f();
  } catch (const std::bad_alloc &) {
    return cdgpu_set_error(CDGPU_ENOMEM, "out of host memory");
  } catch (...) {
    return cdgpu_set_error(CDGPU_ECUDA, "unexpected C++ exception inside libcdgpu");
  }
}
#endif

// ------------------------------------------------------------ device math --
// ProximalBase.shrink: comparison based, NaN -> 0 (oracle/cdref.c: shrink)
__device__ __forceinline__ double cd_shrink(double v, double c) { return v > c ? v - c : (v < -c ? v + c : 0.0); }

// Visit order of RandomIterator mode 1: a keyed bijection of [0, N) (4-round Feistel network on
// 2*hb bits + cycle walking), evaluated per position — no sort, no sequential shuffle.  Same
// definition in oracle/cdref.c:order_perm.  cd_perm maps visit position -> item, cd_perm_inv
// maps item -> visit position.
__host__ __device__ __forceinline__ uint64_t cd_splitmix(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint32_t cd_mix32(uint32_t h) {
  h ^= h >> 16;
  h *= 0x85ebca6bu;
  h ^= h >> 13;
  h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return h;
}
struct PermKey {
  uint32_t key[4];
  uint32_t mask;
  int hb;
  uint32_t N;
};
__host__ __device__ __forceinline__ PermKey cd_perm_key(uint32_t N, uint64_t seed, uint64_t pass) {
  PermKey k;
  uint64_t k0 = cd_splitmix(seed ^ (pass * 0xD1B54A32D192ED03ull)), k1 = cd_splitmix(k0);
  k.key[0] = (uint32_t)k0;
  k.key[1] = (uint32_t)(k0 >> 32);
  k.key[2] = (uint32_t)k1;
  k.key[3] = (uint32_t)(k1 >> 32);
  int bits = 0;
  while (N > 1 && ((N - 1) >> bits) != 0) bits++;
  k.hb = (bits + 1) / 2;
  if (k.hb < 1) k.hb = 1;
  k.mask = (uint32_t)((1ull << k.hb) - 1);
  k.N = N;
  return k;
}
__host__ __device__ __forceinline__ uint32_t cd_perm(const PermKey &k, uint32_t x) {
  do {
    uint32_t L = x >> k.hb, R = x & k.mask;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      uint32_t t = L ^ (cd_mix32(R + k.key[r]) & k.mask);
      L = R;
      R = t;
    }
    x = (L << k.hb) | R;
  } while (x >= k.N);
  return x;
}
__host__ __device__ __forceinline__ uint32_t cd_perm_inv(const PermKey &k, uint32_t x) {
  do {
    uint32_t L = x >> k.hb, R = x & k.mask;
#pragma unroll
    for (int r = 3; r >= 0; --r) {
      uint32_t t = R ^ (cd_mix32(L + k.key[r]) & k.mask);
      R = L;
      L = t;
    }
    x = (L << k.hb) | R;
  } while (x >= k.N);
  return x;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
    v = t < v ? t : v;
  }
  return v;
}

// List order after a pass == closed form of ProximalBase.dropzeros! (the LAST stored entry moves into
// the hole) applied to the reference's pre-compaction list L.  During a full pass the reference's
// `x[k] += b/a` (cd_differentiable_function.jl:102,185,335) APPENDS every visited non-member whose
// tentative value is non-zero, the prox then stores an explicit zero for most of them, and
// dropzeros! at the end of the pass (coordinate_descent.jl:108) removes those again — which
// permutes the survivors: with K survivors, a survivor at position < K stays, and the holes among
// the first K positions are filled, in increasing hole order, by the survivors at positions >= K
// taken in DEcreasing position order.  (Verified against the sequential algorithm, tests/.)
//
// Block-collective over T threads.  Entries e < m_now: coordinate act[e], value val[e], position
// in L = e for e < m_old, newpos[e - m_old] otherwise; an entry survives iff val[e] != 0.
// tmp_i: >= 5*m_now ints, tmp_d: >= m_now doubles, s2: two shared ints.  Result: act/val[0..K),
// K in s2[0]; inlist (nullable) is cleared for dropped coordinates.
template <int T>
__device__ void cd_compact_list(int *act, double *val, int m_old, int m_now, const int *newpos, unsigned char *inlist,
                                int *tmp_i, double *tmp_d, int *s2) {
  // T == 32: warp-collective (several independent warps per CTA may call it), else block-collective
  const int tid = T == 32 ? (int)(threadIdx.x & 31) : (int)threadIdx.x;
  auto sync = [] {
    if (T == 32)
      __syncwarp();
    else
      __syncthreads();
  };
  if (tid == 0) {
    s2[0] = 0;
    s2[1] = 0;
  }
  sync();
  int cnt = 0;
  for (int e = tid; e < m_now; e += T) cnt += (val[e] != 0.0);
  if (cnt) atomicAdd(&s2[0], cnt);
  sync();
  const int K = s2[0];
  if (K == m_now && m_now == m_old) return; // nothing dropped, nothing new: order unchanged
  int *F = tmp_i, *tailE = tmp_i + m_now, *tailP = tmp_i + 2 * m_now, *tailS = tmp_i + 3 * m_now, *ti = tmp_i + 4 * m_now;
  for (int i = tid; i < K; i += T) F[i] = -1;
  sync();
  for (int e = tid; e < m_now; e += T) {
    if (val[e] != 0.0) {
      const int pos = e < m_old ? e : newpos[e - m_old];
      if (pos < K) {
        F[pos] = e;
      } else {
        const int t = atomicAdd(&s2[1], 1);
        tailE[t] = e;
        tailP[t] = pos;
      }
    } else if (inlist) {
      inlist[act[e]] = 0;
    }
  }
  sync();
  const int Tn = s2[1];
  for (int t = tid; t < Tn; t += T) {
    const int pt = tailP[t];
    int r = 0;
    for (int u = 0; u < Tn; ++u) r += tailP[u] > pt;
    tailS[r] = tailE[t];
  }
  sync();
  if (tid < 32) {
    int filled = 0;
    for (int base = 0; base < K; base += 32) {
      const int i = base + tid;
      const bool hole = i < K && F[i] < 0;
      const unsigned b = __ballot_sync(0xffffffffu, hole);
      if (hole) F[i] = tailS[filled + __popc(b & ((1u << tid) - 1u))];
      filled += __popc(b);
    }
  }
  sync();
  for (int i = tid; i < K; i += T) {
    ti[i] = act[F[i]];
    tmp_d[i] = val[F[i]];
  }
  sync();
  for (int i = tid; i < K; i += T) {
    act[i] = ti[i];
    val[i] = tmp_d[i];
  }
  sync();
}

// device-side mirror of cdgpu_stats (host fills device_ms / sigma bookkeeping)
struct DevStats {
  long long passes, full_passes, visits, accepted;
  double maxH;
  int converged, outer_iters;
  double sigma;
};

// number of kernels this library has launched (diagnostic; bench.py's gpu_launches)
extern long long g_cdgpu_launches;
#define CD_COUNT_LAUNCH(k) (g_cdgpu_launches += (k))

struct CovResume;
// ---------------------------------------------------------------- handle --
struct cdgpu_handle_s {
  int kind = -1, device = 0;
  int64_t n = 0, p = 0, ld = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // naive: X (n x p, ld), y, w, r=state(n), colsq (p) ; quad: A (p x p, ld), b, state=Ax(p), ainv (p)
  double *dX = nullptr, *dy = nullptr, *dw = nullptr, *dstate = nullptr, *daux = nullptr;
  bool ownX = false, owny = false, ownw = false;
  unsigned char *dcommon = nullptr; // one allocation behind the iterate + scratch pointers below
  // iterate
  double *dbeta = nullptr;          // dense p
  int *dact = nullptr;              // active list (0-based coordinates), capacity p
  double *dactval = nullptr;        // values in list order, capacity p
  int *dnact = nullptr;             // 1 int
  unsigned char *dinlist = nullptr; // p flags
  double *domega = nullptr;         // p (scratch copy of the caller's weights)
  // scratch for the sweep kernels
  double *dscr = nullptr;           // 12 * p + 8 * n doubles
  int *discr = nullptr;             // 8 * p ints
  unsigned char *dbscr = nullptr;   // 2 * p bytes
  double *dgram = nullptr;          // naive handles: active-set Gram scratch (allocated at first solve)
  void *dtiles = nullptr;           // Gram tile schedule (int2[ntiles])
  int ntiles = 0;
  DevStats *dstats = nullptr;       // path stats, grown on demand
  int64_t nstats = 0;
  double *dlam = nullptr;
  int64_t nlam = 0;
  // path output
  long long *dcolptr = nullptr, *drowval = nullptr;
  double *dnzval = nullptr;
  int64_t outcap = 0, outcols = 0;
  int *dflag = nullptr;             // device status word(s)
  double *drsnap = nullptr;         // NaiveArgs::rsnap
  double *dchain = nullptr;         // scratch of the team chain engine (chain_engine.cuh: Multi::hpass, Multi::seq)
  double *dtall = nullptr;          // per-CTA partial sums of the row-distributed sqrt-lasso sweep (tall_sweep.cu)
  double gram_ms = 0.0;
  int sm_count = 0, max_cluster = 0;
  // lazy covariance form (lazy_gram.cu): dX is then the column CACHE (ld x lz_cap), columns found through dslot
  bool lazy = false;
  const double *lzX = nullptr;      // the data, n x p (ld lz_ldx) on the device
  const double *lzy = nullptr;
  const double *lzw = nullptr;      // optional observation weights: A = X'WX/n, b = -X'Wy/n (tall CDWeightedLSLoss)
  cdgpu_handle_s *tall = nullptr;   // naive LS/WLS handle too tall for the residual-in-shared-memory kernel: the solves run
                                    // on this inner lazy covariance handle over the same X, y (w) — same minimiser
  bool own_lzX = false, own_lzy = false;
  int64_t lz_n = 0, lz_ldx = 0;
  int lz_cap = 0, lz_used = 0;      // slots of the cache / slots filled
  int *dslot = nullptr;             // [p] coordinate -> slot, -1: column not formed
  double *ddiag = nullptr;          // [p] diag(A)
  double *dgather = nullptr;        // [lz_ldb x 128] the gathered columns of one batch (second GEMM operand)
  int64_t lz_ldb = 0;
  CovResume *dresume = nullptr;
  int *dbatch = nullptr;            // [128] coordinates of the batch being formed
  // speculative batch formed on a second stream WHILE the sweep kernel runs (the cluster kernel occupies 16 SMs)
  cudaStream_t lz_stream2 = nullptr;
  cudaEvent_t lz_spec_ev = nullptr;
  double *dgather2 = nullptr;
  int *dbatch2 = nullptr;
  int spec_cols[128];               // host copy of the speculative batch
  int spec_n = 0, spec_slot0 = 0;
  bool spec_inflight = false;
  int next_cand[256];               // runner-up candidates of the last scoring (best first)
  int next_n = 0;
  int *hslot = nullptr;             // host mirror of dslot
  cudaEvent_t lz_ev0 = nullptr, lz_ev1 = nullptr;
  int lz_batches = 0, lz_pauses = 0; // statistics of the last solve
  long long lz_form_cols = 0;        // columns formed by the blocking (timed) batches of the last solve
  double lz_form_ms = 0.0;           // device time spent forming columns during the last solve
  double sweep_ms = 0.0;             // device time of the covariance sweep-kernel launches of the last solve
  bool sweep_pending = false;
  cudaEvent_t sw_ev0 = nullptr, sw_ev1 = nullptr;
};

// device selection (+ memory-pool release threshold) and the per-device free list of stream/event sets (api.cu)
int cd_use_device(int device);
struct StreamSet {
  cudaStream_t stream;
  cudaEvent_t ev0, ev1;
};
int stream_set_acquire(int device, StreamSet *out);
void stream_set_release(int device, const StreamSet &s);

// State of a covariance-form path that stopped because a coordinate entered whose column of A = X'X/n has not been
// formed yet (lazy covariance handle, api.cu).  The kernel leaves through a consistent point INSIDE the full pass
// (before the entering step is applied); the host forms the missing column(s) and launches again with valid = 1.
struct CovResume {
  int valid;        // host -> kernel: continue from this record
  int need_k;       // kernel -> host: the coordinate whose column is missing
  int li;           // lambda index
  int conv, m_bound, m_old, tzflag, pad0;
  long long iter, out_off, cols_done, curpos;
  unsigned long long pass_counter;
  double maxH;      // of the pass in progress
  DevStats st;
};

// ------------------------------------------------------------- launchers --
struct CovArgs {
  const double *A;
  long long lda;
  int p;
  const double *b, *ainv, *omega; // omega may be null
  double *Ax, *beta;
  int *act;
  double *actval;
  int *nact;
  unsigned char *inlist;
  double *scr;  // >= 8p doubles
  int *iscr;    // >= 8p ints
  unsigned char *bscr; // >= 2p bytes
  const double *lambdas;
  int nlambda;
  int accumulate; // 1: all lambdas are one solve (cold-start continuation): stats summed into stats[0], no path output
  long long maxIter;
  double optTol;
  int randomize;
  unsigned long long seed;
  long long max_hat_s; // <0: none
  long long *colptr, *rowval;
  double *nzval;
  long long capacity;
  int *flag; // [0]=status (0 ok, 1 capacity, 2 active set too large, 3 paused: column resume->need_k missing), [1]=columns done
  DevStats *stats;
  long long *prof; // optional [8]: SM cycles spent per phase by CTA 0 (CDGPU_PROFILE=1)
  // lazily formed columns: column k of the p x p matrix lives at A + colslot[k]*lda (colslot[k] < 0: not formed yet);
  // null for a fully formed A.  resume: see CovResume (null: never pause; a missing column is then an error)
  const int *colslot;
  CovResume *resume;
  int events_only; // diagnostics: 1 = the event-by-event full pass of round 1 instead of chain + verify
  double *chain_scr; // CD_MULTI_SCR_BYTES of global scratch for the team chain engine
};
int launch_cov_path(cdgpu_handle_s *h, const CovArgs &a);
int launch_cov_init(cdgpu_handle_s *h, const double *A, long long lda, int p, const int *act, const double *actval,
                    const int *nact, double *Ax, double *beta, unsigned char *inlist, const int *colslot);
int launch_lambda_max_quad(cdgpu_handle_s *h, const double *b, const double *omega, int p, double *out);
int launch_extract_ainv(cdgpu_handle_s *h, const double *A, long long lda, int p, double *ainv);
int launch_check_symmetric(cdgpu_handle_s *h, const double *A, long long lda, int p, int *flag);
int launch_diag_sqrt(cdgpu_handle_s *h, const double *A, long long lda, int p, double *out);

// Gram (gram_dmma.cu): G = X'X / n_total (lower tiles computed, mirrored), c = -X'y / n_total
// mode 0: raw sums, 1: sums / divisor, 2: accumulate raw sums into G | c, 3: accumulate, then divide
int launch_gram(cdgpu_handle_s *h, const double *X, long long n, int p, long long ldx, const double *y, double *G,
                double *c, double divisor, int mode);
int launch_scale_gram(cdgpu_handle_s *h, double *G, double *c, int p, double n_total);
// lazy_gram.cu
int launch_diag_xty(cdgpu_handle_s *h, const double *X, long long n, int p, long long ldx, const double *y, const double *w,
                    double divisor, double *diag, double *b, double *ainv, int accumulate, int finish);
int launch_lazy_score(cdgpu_handle_s *h, const double *Ax, const double *b, const double *omega, const int *slot, int p,
                      double *out);
int launch_gather_cols(cudaStream_t stream, const double *X, long long ldx, long long n, const double *w, const int *idx, int nb,
                       int nbpad, double *B, long long ldb, int *slot, int slot0); // slot == null: no slot assignment
int launch_assign_slots(cudaStream_t stream, const int *idx, int nb, int slot0, int *slot);
int launch_fill_int(cdgpu_handle_s *h, int *a, int n, int v);
int launch_sqrt_vec(cdgpu_handle_s *h, const double *a, int n, double *out);
// refit.cu: least squares on a support (refitLassoPath); scratch >= ld*ns + ns doubles, ld = ns rounded up to even
int launch_refit(cdgpu_handle_s *h, const int *dS, int ns, double *scratch, int *flag);
int launch_gemm_tn(cudaStream_t stream, int sm_count, const double *A, int pa, long long lda, const double *B, int pb,
                   long long ldb, long long n, double *C, long long ldc, double divisor, void **tiles_out);

int launch_gemm_tn_split(cudaStream_t stream, int sm_count, const double *A, int pa, long long lda, const double *B, int pb,
                         long long ldb, long long n, double *C, long long ldc, double divisor);

// naive sweeps (naive_sweep.cu)
struct NaiveArgs {
  int kind;
  const double *X;
  long long ldx;
  int n, p;
  const double *y, *w, *colsq, *omega;
  double *r, *beta;
  int *act;
  double *actval;
  int *nact;
  unsigned char *inlist;
  double *scr;
  int *iscr;
  unsigned char *bscr;
  const double *lambdas;
  int nlambda;
  int accumulate;
  long long maxIter;
  double optTol;
  int randomize;
  unsigned long long seed;
  long long max_hat_s;
  long long *colptr, *rowval;
  double *nzval;
  long long capacity;
  int *flag;
  DevStats *stats;
  // scaled lasso (sigma loop on device): lambdas[0] = lambda, sigma0 = initial sigma
  int scaled;
  long long outerMaxIter;
  double outerTol, sigma0;
  long long *prof; // optional [10]: SM cycles per phase on CTA 0 (CDGPU_PROFILE=1)
  double *gram;    // scratch for the covariance-form active engine: gram_cap^2 + gram_cap doubles (or null)
  int gram_cap;    // largest active set the scratch holds (cd_gram_cap(p))
  int multi_ok;    // grid-distributed chain engine for large active sets (CDGPU_NAIVE_MULTI=0 disables)
  int pipeline;    // split-phase rounds of the full pass (CDGPU_NAIVE_PIPELINE=0 disables)
  int plan;        // members' steps of a full pass planned by one chain pass (CDGPU_NAIVE_PLAN=0 disables)
  double *chain_scr; // CD_MULTI_SCR_BYTES of global scratch for the team chain engine
  double *rsnap;   // 2n doubles: r at the start of a planned super-window of the full pass (two buffers)
  int dense;       // dense mode of the full pass: segments planned over members + candidates (CDGPU_NAIVE_DENSE=0 disables)
  int replan;      // members still to come are planned again after an entering coordinate moved (-DCDGPU_WITH_REPLAN builds, CDGPU_NAIVE_REPLAN=1)
};
// offset (doubles, even) of the tail of a handle's scratch: 16p doubles for the result buffers of the full-pass rounds
constexpr int CD_GCAP = 4096; // largest active set of a naive handle that runs on the chain engines (Gram scratch 134 MB)
// global scratch of the team chain engine (chain_engine.cuh: Multi::gT, Multi::hT: CD_GCAP tagged values of 16 bytes each)
constexpr size_t CD_MULTI_SCR_BYTES = (size_t)2 * CD_GCAP * 16;
inline size_t cd_gram_cap(size_t p) { return p < 2048 ? 2048 : (p < (size_t)CD_GCAP ? ((p + 1) & ~(size_t)1) : (size_t)CD_GCAP); } // even
inline size_t cd_scr_tail(size_t p, size_t n) { return (15 * p + 8 * n + 64 + 4 * (size_t)CD_GCAP + 64 + 1) & ~(size_t)1; }
int launch_naive_path(cdgpu_handle_s *h, const NaiveArgs &a);
int launch_tall_sqrt(cdgpu_handle_s *h, const NaiveArgs &a); // tall_sweep.cu: rows dealt over the grid (n of any size)
bool naive_fits(long long n, bool has_w);
int launch_naive_init(cdgpu_handle_s *h, const NaiveArgs &a); // r = y - X beta, beta dense, inlist
int launch_colsq(cdgpu_handle_s *h, const double *X, long long ldx, int n, int p, const double *w, double *out,
                 bool sqrt_over_n);
int launch_abs_xty(cdgpu_handle_s *h, const double *X, long long ldx, int n, int p, const double *y, double *out);
int launch_lambda_max_naive(cdgpu_handle_s *h, int kind, const double *X, long long ldx, int n, int p, const double *y,
                            const double *w, const double *omega, double *scr, double *out);
