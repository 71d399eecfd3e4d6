// naive_sweep.cu — naive/residual-form active-set coordinate descent for CDLeastSquaresLoss,
// CDWeightedLSLoss and CDSqrtLassoLoss (src/cd_differentiable_function.jl:43-291) driven by
// _coordinateDescent!/_cdPass! (src/coordinate_descent.jl:65-110), with the scaledLasso! sigma loop
// (src/lasso.jl:131-141) on the device.  One persistent cooperative kernel per solve / path.
//
// B200 design:
//  * every CTA keeps a private copy of the residual r (and the weights w) in shared memory; the
//    same updates are applied to every copy, so the copies stay bit-identical with no traffic.
//  * FULL pass = exact Gauss-Seidel by speculation over chunks of columns: one warp per column
//    streams X[:,k] from HBM once (8n bytes/visit, the algorithmic minimum), forms X_k'r against the
//    shared r, evaluates the closed-form update and publishes h.  One grid barrier later every CTA
//    finds the FIRST coordinate of the chunk that moves, applies r -= X_k h to its copy (the column
//    is L2 resident) and only the columns behind it are re-evaluated (from L2, not HBM).
//  * the column norms a_k = sum_i [w_i] X_ik^2, which the reference recomputes on every visit
//    (:96-99), are formed once per handle.
//  * ACTIVE-SET passes are a sequential chain over a few columns: CTA 0 runs them alone against its
//    shared r (column reads are L2 hits, next column prefetched), then publishes r.
//  * sqrt-lasso: s, ||r+||^2 of :254-266 follow from d = X_k'r, ||r||^2 and a_k
//    (s = d + a_k x_k, ||r+||^2 = ||r||^2 + 2 x_k d + x_k^2 a_k), so a visit is still one column read.
#include <cooperative_groups.h>

#include "chain_engine.cuh"
#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int NV_T = 512;
constexpr int NV_W = NV_T / 32;
constexpr int NV_GCAP_ = CD_GCAP; // == NV_GCAP below: stride of the per-entry scratch arrays in global memory

struct __align__(16) HEntry {
  double h, nw;
  int app, pad[3]; // app == 0: a non-member whose tentative value was exactly zero (not appended by setindex!)
};
static_assert(sizeof(HEntry) == 32, "HEntry");

struct NBcast {
  long long npasses, visits, accepted;
  double maxH;
  int conv, nact, status, pad;
};

struct NSmem {
  double red[2][NV_W];
  unsigned int redu[NV_W];
  double bval[2], bval2[2];
  int nact, flag, nonapp;
  int s2[2];
  unsigned warr; // warp arrivals at the warp-granular round barrier of the full pass (monotonic)
  chain::Shared ch;
};

struct NCtx {
  const NaiveArgs &a;
  cg::grid_group &grid;
  NSmem *sm;
  double *r, *w; // shared
  double rr;     // ||r||^2 (every thread of every CTA holds the same value)
  HEntry *hbuf;  // global, 4 * CH (slot = round & 3)
  NBcast *bc;    // global
  // shared state of the covariance-form active engine (chain_engine.cuh), gcap entries (0: engine disabled)
  int *e_row, *e_coord;
  double *e_g, *e_be, *e_stage;
  unsigned short *e_ord, *e_pos;
  int gcap;
  int CH, G, bid;
  unsigned bar_target; // grid barriers so far, times G (every thread keeps the same value)
  unsigned rnd;        // rounds of the full pass so far (slot of the first-mover word / result buffer = rnd & 3)
};

__device__ __forceinline__ double block_sum(NSmem *sm, double v, int slot) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sm->red[slot][threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int i = 0; i < NV_W; ++i) t += sm->red[slot][i];
  return t;
}
__device__ __forceinline__ void block_sum2(NSmem *sm, double &u, double &v) {
  u = warp_sum(u);
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) {
    sm->red[0][threadIdx.x >> 5] = u;
    sm->red[1][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  double t0 = 0.0, t1 = 0.0;
#pragma unroll
  for (int i = 0; i < NV_W; ++i) {
    t0 += sm->red[0][i];
    t1 += sm->red[1][i];
  }
  u = t0;
  v = t1;
}

// closed-form coordinate update given d = X_k' (w .* r):
//   LS/WLS  cd_differentiable_function.jl:101-104 / :184-187;  sqrt  :271-283
__device__ __forceinline__ void coord_update(int kind, int n, double d, double a, double old, double lam, double om,
                                             double rr, double &nw, double &h, bool &tentative_nz) {
  if (kind == CDGPU_LOSS_SQRT) {
    const double s = d + a * old;
    const double rsq = rr + 2.0 * old * d + old * old * a;
    const double l = lam * om;
    const double t = l * sqrt(rsq);
    if (fabs(s) <= t)
      nw = 0.0;
    else if (s > t)
      nw = (s - l / sqrt(1.0 - l * l / a) * sqrt(rsq - s * s / a)) / a;
    else
      nw = (s + l / sqrt(1.0 - l * l / a) * sqrt(rsq - s * s / a)) / a;
    tentative_nz = nw != 0.0; // x[k] = newVal: appended only when non-zero (:278-283)
  } else {
    const double v = __dadd_rn(old, d / a);
    tentative_nz = v != 0.0; // x[k] += b/a appends whenever the sum is non-zero (:102,:185)
    const double thr = __dmul_rn(__dmul_rn((double)n / a, lam), om);
    nw = cd_shrink(v, thr);
  }
  h = nw - old;
}

// streaming 16-byte load that does not pollute L1 (each column is read once per visit)
__device__ __forceinline__ double2 ldg_stream2(const double2 *p) {
  double2 v;
  // (a weak load, not .nc: the warp barriers that pin the issue order in warp_col_dot_t do not order read-only loads)
#ifdef CDGPU_STREAM_NC
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
#else
  asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
#endif
  return v;
}

// one warp: d = sum_i X[i,k] r_i [w_i].  16-byte streaming loads, software pipelined in half-batches of four (2 KB per
// warp): the next four loads are issued before the last four are consumed, so a warp has 2-4 KB in flight at every
// moment whatever order the compiler gives the instructions (left alone it splits a batch of eight loads around the
// first multiply-adds and the warp drains to zero loads in flight between batches: the streaming rounds of a full
// pass ran 10 % slower after an unrelated change elsewhere in the kernel); a warp barrier behind each group of loads
// pins that order for the assembler (weak loads: read-only ones may cross it).  The tail (<= 8 elements per lane) is
// one batch of clamped loads, masked.  Summation order: full batches alternate (s0,s1)/(s2,s3) over u = 0..7, the tail
// adds to (s0,s1).  (Not a function call: a call in this kernel makes r a generic pointer and
// costs every phase 60 %, measured.)
template <bool HASW>
__device__ __forceinline__ double warp_col_dot_t(const NCtx &c, const double *col) {
  const int n = c.a.n, lane = threadIdx.x & 31;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  if ((reinterpret_cast<uintptr_t>(col) & 15) == 0) {
    const double2 *c2 = reinterpret_cast<const double2 *>(col);
    const double2 *r2 = reinterpret_cast<const double2 *>(c.r);
    const double2 *w2 = reinterpret_cast<const double2 *>(c.w);
    const int np = n >> 1;
    int i = lane;
    auto consume4 = [&](double2(&x)[4], int at) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const double2 rv = r2[at + 32 * u];
        if (HASW) {
          const double2 wv = w2[at + 32 * u];
          x[u].x *= wv.x;
          x[u].y *= wv.y;
        }
        if (u & 1) {
          s2 = fma(x[u].x, rv.x, s2);
          s3 = fma(x[u].y, rv.y, s3);
        } else {
          s0 = fma(x[u].x, rv.x, s0);
          s1 = fma(x[u].y, rv.y, s1);
        }
      }
    };
    const int nb = np >> 8; // batches of 256 double2 that every lane takes in full (warp-uniform: the barriers below)
    if (nb > 0) {
      double2 xa[4], xb[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) xa[u] = ldg_stream2(c2 + i + 32 * u);
      for (int bt = 0; bt < nb; ++bt) {
#pragma unroll
        for (int u = 0; u < 4; ++u) xb[u] = ldg_stream2(c2 + i + 128 + 32 * u);
        __syncwarp(); // scheduling fence: the four loads above are issued before anything below
        consume4(xa, i);
        // the first half of the next batch; behind the last batch the same addresses once more (an L1/L2 hit, never
        // used): a conditional load would be a predicated load into temporaries that are copied at once — a stall for
        // the full memory latency right behind the issue
        const int nx = bt + 1 < nb ? i + 256 : i;
#pragma unroll
        for (int u = 0; u < 4; ++u) xa[u] = ldg_stream2(c2 + nx + 32 * u);
        __syncwarp();
        consume4(xb, i + 128);
        i += 256;
      }
    }
    if (i < np) { // tail: at most 8 elements per lane
      double2 xt[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) xt[u] = ldg_stream2(c2 + min(i + 32 * u, np - 1));
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = i + 32 * u;
        if (j < np) {
          const double2 rv = r2[j];
          if (HASW) {
            xt[u].x *= w2[j].x;
            xt[u].y *= w2[j].y;
          }
          s0 = fma(xt[u].x, rv.x, s0);
          s1 = fma(xt[u].y, rv.y, s1);
        }
      }
    }
    if ((n & 1) && lane == 0) s2 = fma(HASW ? col[n - 1] * c.w[n - 1] : col[n - 1], c.r[n - 1], s2);
  } else {
    for (int i = lane; i < n; i += 32) s0 = fma(HASW ? __ldg(col + i) * c.w[i] : __ldg(col + i), c.r[i], s0);
  }
  return warp_sum((s0 + s1) + (s2 + s3));
}
__device__ __forceinline__ double warp_col_dot(const NCtx &c, const double *col) {
  return c.w ? warp_col_dot_t<true>(c, col) : warp_col_dot_t<false>(c, col);
}

// Barrier of the whole cooperative grid: one atomic arrive + acquire-poll on a counter in global memory (the
// cooperative launch guarantees co-residency).  Every CTA must call it the same number of times; cheaper than
// cg::grid_group::sync() over 148 CTAs.  The counter is zeroed at kernel start.
__device__ __forceinline__ void fast_grid_sync(NCtx &c) {
  unsigned *ctr = reinterpret_cast<unsigned *>(c.a.flag + 7);
  c.bar_target += (unsigned)c.G;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    unsigned v;
    int spins = 0;
    for (;;) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if (v >= c.bar_target) break;
      if (++spins > 64) __nanosleep(256); // a long wait (e.g. CTAs outside the 16-CTA engine team): stop hammering L2
    }
  }
  __syncthreads();
}

// r -= X_k * h on this CTA's copy; refreshes ||r||^2 when the loss needs it
__device__ __forceinline__ void apply_step(NCtx &c, const double *col, double h) {
  const int n = c.a.n;
  if (c.a.kind == CDGPU_LOSS_SQRT) {
    double acc = 0.0;
    for (int i0 = threadIdx.x; i0 < n; i0 += 8 * NV_T) {
      double xv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) xv[u] = (i0 + u * NV_T < n) ? __ldg(col + i0 + u * NV_T) : 0.0;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * NV_T;
        if (i < n) {
          const double v = __dsub_rn(c.r[i], __dmul_rn(xv[u], h));
          c.r[i] = v;
          acc = fma(v, v, acc);
        }
      }
    }
    c.rr = block_sum(c.sm, acc, 0);
    __syncthreads();
  } else {
    for (int i0 = threadIdx.x; i0 < n; i0 += 8 * NV_T) {
      double xv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) xv[u] = (i0 + u * NV_T < n) ? __ldg(col + i0 + u * NV_T) : 0.0;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * NV_T;
        if (i < n) c.r[i] = __dsub_rn(c.r[i], __dmul_rn(xv[u], h));
      }
    }
    __syncthreads();
  }
}

// r -= sum_t X[:, k_t] h_t over the planned entries t in [t0, t1) (increasing visit positions), element by element in
// visit order with the same non-fused operations as one apply_step per entry; ||r||^2 refreshed as apply_step does.
// Touches this CTA's r only: the iterate and the list are written by commit_planned once the steps are final.
__device__ void apply_planned_r(NCtx &c, int t0, int t1) {
  const NaiveArgs &a = c.a;
  const int n = a.n, tid = threadIdx.x;
  for (int i0 = tid; i0 < n; i0 += 2 * NV_T) {
    const int i1 = i0 + NV_T;
    double v0 = c.r[i0], v1 = i1 < n ? c.r[i1] : 0.0;
    int t = t0;
    for (; t + 4 <= t1; t += 4) { // 8 independent loads in flight, then the dependent updates in visit order
      double x0[4], x1[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const double *col = a.X + (long long)c.e_coord[t + u] * a.ldx;
        x0[u] = __ldg(col + i0);
        x1[u] = i1 < n ? __ldg(col + i1) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const double h = c.e_g[t + u];
        v0 = __dsub_rn(v0, __dmul_rn(x0[u], h));
        v1 = __dsub_rn(v1, __dmul_rn(x1[u], h));
      }
    }
    for (; t < t1; ++t) {
      const double *col = a.X + (long long)c.e_coord[t] * a.ldx;
      const double h = c.e_g[t];
      v0 = __dsub_rn(v0, __dmul_rn(__ldg(col + i0), h));
      if (i1 < n) v1 = __dsub_rn(v1, __dmul_rn(__ldg(col + i1), h));
    }
    c.r[i0] = v0;
    if (i1 < n) c.r[i1] = v1;
  }
  __syncthreads();
  if (a.kind == CDGPU_LOSS_SQRT) { // same rows per thread and same summation order as apply_step
    double acc = 0.0;
    for (int i0 = tid; i0 < n; i0 += 8 * NV_T) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * NV_T;
        if (i < n) acc = fma(c.r[i], c.r[i], acc);
      }
    }
    c.rr = block_sum(c.sm, acc, 0);
    __syncthreads();
  }
}
// the planned entries [t0, t1) are final: statistics (every CTA), the dense iterate and the list (CTA 0)
__device__ void commit_planned(NCtx &c, int t0, int t1, double &maxH, long long &accepted) {
  const NaiveArgs &a = c.a;
  const int tid = threadIdx.x;
  for (int t = t0; t < t1; ++t) {
    const double h = c.e_g[t];
    maxH = fmax(maxH, fabs(h));
    accepted += h != 0.0;
  }
  if (c.bid == 0) {
    for (int t = t0 + tid; t < t1; t += NV_T)
      if (c.e_g[t] != 0.0) __stcg(a.beta + c.e_coord[t], c.e_be[t]);
    if (tid == 0) // dense mode plans candidates too: a non-member that moves is appended (setindex!), in visit order
      for (int t = t0; t < t1; ++t) {
        const int k = c.e_coord[t];
        if (c.e_g[t] != 0.0 && !a.inlist[k]) {
          __stcg(a.inlist + k, (unsigned char)1);
          a.act[c.sm->nact] = k;
          c.sm->nact += 1;
        }
      }
  }
}

// plan arrays in global memory (CTA 0 writes, everybody copies them to shared memory)
struct NPlan {
  int *k, *pos, *row; // coordinate, visit position, list index (= row of the active Gram) of the t-th visited member
  int *ext;           // dense mode: the coordinates planned for a segment of a full pass (members + candidates), NV_PLAN_MAX
  double *h, *nw;     // (row[512] = leading dimension of G, row[513] = entries in ext, row[514] = ext was truncated)
};
__device__ __forceinline__ NPlan plan_arrays(const NaiveArgs &a) {
  NPlan P;
  P.k = a.iscr + 10 * (long long)a.p + 64; // 3 x NV_PLAN_MAX ints behind the sweep's scratch ints (handle_common_alloc)
  P.pos = P.k + 512;
  P.row = P.pos + 512;
  P.ext = P.row + 520;
  P.h = a.scr + 8 + 9 * (long long)a.p + 32 + 2 * NV_GCAP_;
  P.nw = P.h + NV_GCAP_;
  return P;
}

// ------------------------------------------------------------------ full pass --
// A ROUND evaluates a window of columns against the same r, then one grid barrier makes the first mover of the window
// known to every CTA.  Rounds are numbered over the whole kernel (c.rnd); round s owns slot s & 3 of the result buffers
// and of the first-mover words; every round has exactly one barrier (arrive_s after check_{s-1}).
//  * CTA-per-column rounds (right behind a mover: movers cluster, lowest latency) use the CTA-wide barrier.
//  * Warp-per-column rounds (streaming regime) are WARP-GRANULAR and SPLIT-PHASE: a warp that has finished its columns
//    of round s arrives (shared counter; the last warp of the CTA arrives at the grid counter) and, once two rounds in a
//    row were clean, goes straight on to the columns of round s+1 — evaluated against the same r on the assumption that
//    round s is clean too — and only then waits for barrier s.  In the clean streaming regime no SM ever drains its
//    memory pipeline at a barrier.  A mover in round s discards round s+1 (its barrier still runs, empty).
//  * first-mover word of round s is reset by CTA 0 at check_{s+1}: everyone read it at check_s, i.e. before arriving at
//    barrier s+1, and its next writers (round s+4) start behind barrier s+2, which CTA 0 joins after that reset.
__device__ __forceinline__ void round_arrive_warp(NCtx &c) {
  unsigned *ctr = reinterpret_cast<unsigned *>(c.a.flag + 7);
  c.bar_target += (unsigned)c.G;
  __syncwarp();
  if ((threadIdx.x & 31) == 0) {
    __threadfence(); // this warp's results of the round
    const unsigned old = atomicAdd(&c.sm->warr, 1u);
    if ((old % NV_W) == NV_W - 1) {
      __threadfence();
      atomicAdd(ctr, 1u);
    }
  }
}
__device__ __forceinline__ void round_wait_warp(NCtx &c) {
  unsigned *ctr = reinterpret_cast<unsigned *>(c.a.flag + 7);
  if ((threadIdx.x & 31) == 0) {
    unsigned v;
    int spins = 0;
    for (;;) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if (v >= c.bar_target) break;
      if (++spins > 8) __nanosleep(128);
    }
  }
  __syncwarp();
}
__device__ __forceinline__ void round_arrive_cta(NCtx &c) {
  unsigned *ctr = reinterpret_cast<unsigned *>(c.a.flag + 7);
  c.bar_target += (unsigned)c.G;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
  }
}
__device__ __forceinline__ void round_wait_cta(NCtx &c) {
  unsigned *ctr = reinterpret_cast<unsigned *>(c.a.flag + 7);
  if (threadIdx.x == 0) {
    unsigned v;
    int spins = 0;
    for (;;) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if (v >= c.bar_target) break;
      if (++spins > 64) __nanosleep(256);
    }
  }
  __syncthreads();
}

constexpr int NV_PLAN_MAX = 512;  // largest list whose steps are planned ahead of a full pass (member_plan)
constexpr int NV_REPLAN_MIN = 8;  // a non-member moved: the remaining members are planned again when at least this many are left
constexpr int NV_REPLAN_MAX = 8;  // ... at most this often per pass (each costs two grid barriers and a chain over the rest)
static_assert(NV_PLAN_MAX <= NV_T, "replan_members handles one entry per thread");
__device__ void replan_members(NCtx &c, double lam, int t0, int t1);

struct NWin {
  int q0, qlen, W, slot;
  bool cta;
};

// evaluate the columns of a window against the current r: tentative (h, new value) per position, first mover by atomicMin
__device__ __forceinline__ void eval_window(NCtx &c, const NWin &wn, double lam, const PermKey &pk, bool ordered,
                                            unsigned int *words, int *nonapp_flags, int ja, int jb, bool warp_mode) {
  const NaiveArgs &a = c.a;
  NSmem *sm = c.sm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  HEntry *hb = c.hbuf + (size_t)wn.slot * c.CH;
  unsigned int *gmin = words + wn.slot;
  int *nonapp_flag = nonapp_flags + wn.slot;
  if (wn.cta && !warp_mode) { // CTA-per-column mode
    const int j = c.bid;
    if (j < wn.qlen) {
      const int k = ordered ? wn.q0 + j : (int)cd_perm(pk, (uint32_t)(wn.q0 + j));
      const double *col = a.X + (long long)k * a.ldx;
      double s = 0.0;
      for (int t0 = tid; t0 < a.n; t0 += 8 * NV_T) {
        double xv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) xv[u] = (t0 + u * NV_T < a.n) ? __ldg(col + t0 + u * NV_T) : 0.0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int t = t0 + u * NV_T;
          if (t < a.n) s = fma(c.w ? xv[u] * c.w[t] : xv[u], c.r[t], s);
        }
      }
      const double d = block_sum(sm, s, 1);
      if (tid == 0) {
        double nw, h;
        bool tnz;
        coord_update(a.kind, a.n, d, __ldg(a.colsq + k), __ldcg(a.beta + k), lam, a.omega ? __ldg(a.omega + k) : 1.0,
                     c.rr, nw, h, tnz);
        const int app = (a.kind == CDGPU_LOSS_SQRT || tnz || __ldcg(a.inlist + k)) ? 1 : 0;
        __stcg(reinterpret_cast<double2 *>(hb + j), make_double2(h, nw));
        __stcg(&hb[j].app, app);
        if (!app) __stcg(nonapp_flag, 1);
        if (h != 0.0) atomicMin(gmin, (unsigned int)j);
      }
    }
    return;
  }
  // position j of the window belongs to CTA j % G, warp (j / G) % NV_W; [ja, jb): the part of the window evaluated now
  const int stride = c.G * NV_W, jw = c.bid + c.G * warp;
  for (int j = ja <= jw ? jw : jw + (ja - jw + stride - 1) / stride * stride; j < jb; j += stride) {
    const int k = ordered ? wn.q0 + j : (int)cd_perm(pk, (uint32_t)(wn.q0 + j));
    const double d = warp_col_dot(c, a.X + (long long)k * a.ldx);
    if (lane == 0) {
      double nw, h;
      bool tnz;
      coord_update(a.kind, a.n, d, __ldg(a.colsq + k), __ldcg(a.beta + k), lam, a.omega ? __ldg(a.omega + k) : 1.0,
                   c.rr, nw, h, tnz);
      // sqrt-lasso never appends temporarily (x[k] = newVal, :278-283), so nothing to track there
      const int app = (a.kind == CDGPU_LOSS_SQRT || tnz || __ldcg(a.inlist + k)) ? 1 : 0;
      __stcg(reinterpret_cast<double2 *>(hb + j), make_double2(h, nw));
      __stcg(&hb[j].app, app);
      if (!app) __stcg(nonapp_flag, 1);
      if (h != 0.0) atomicMin(gmin, (unsigned int)j); // first position of the round that moves
    }
  }
}

// PLANNED MEMBERS (mP > 0): the members of the iterate almost always move in a full pass (an active phase only converges
// to optTol), and each of them used to cost a round of its own: evaluation of the window behind it, grid barrier, apply.
// With a plan (member_plan below: the members' steps, computed beforehand by one chain pass over the active Gram in
// VISIT order on the assumption that no non-member moves) every CTA applies a member's step to its copy of r when the
// sweep reaches its position — no barrier, consecutive members in one fused update — and windows hold non-members only
// and never cross a planned position.  A non-member that does move voids the rest of the plan: the remaining members
// are then evaluated like any other column.  c.e_coord / e_row / e_g / e_be (idle during a full pass) hold the plan:
// coordinate, visit position, h, new value, in visit order.
// DENSE MODE (mover-dense passes: cold starts, small penalties).  A pass is cut into SEGMENTS.  Before a segment the grid
// evaluates every remaining column once against the current r (dense_prescan) and the plan holds, next to the members,
// the CANDIDATES: non-members that would move at 0.9 lambda; one chain pass over the Gram of members + candidates gives
// all their steps (a candidate that does not move after all gets h = 0 and is not appended).  The segment ends — and the
// rest of the pass is planned again from the position reached (`q_resume` < p) — when a coordinate outside the plan moves
// and enough of the plan is left, or when a truncated plan (more than NV_PLAN_MAX entries) is used up.  Same iterates
// and list order as the one-round-per-mover pass: planned or not, every coordinate is updated at its visit position
// with the step the sequential algorithm takes there.
constexpr int NV_DENSE_MIN_LEFT = 12; // planned entries left that make a new plan cheaper than a round each
constexpr int NV_DENSE_EVENTS = 4;    // ... or this many unplanned movers since the segment started
// (DENSE is a template parameter of the path kernel: problems whose X stays in L2 get the instance with dense mode and
// super-windows; everything else gets the instance without that code — merely compiling it into the kernel cost the
// HBM-bound streaming rounds 10 %, at C3 1.44 -> 1.59 ms, through register allocation.)
template <bool DENSE>
__device__ double full_pass(NCtx &c, double lam, unsigned long long pass_counter, long long &accepted, long long *pf,
                            int nact_hint, int mP, int q_start, bool dense, bool truncated, int &q_resume, int &events,
                            int &plan_used, int &plan_moved) {
  const NaiveArgs &a = c.a;
  NSmem *sm = c.sm;
  const int tid = threadIdx.x;
  const bool ordered = a.randomize == 0;
  const PermKey pk = cd_perm_key((uint32_t)a.p, a.seed, pass_counter);
  unsigned int *words = reinterpret_cast<unsigned int *>(a.flag + 8); // 4 rotating words: first mover of a round
  int *nonapp_flags = a.flag + 12; // per slot: some warp produced a non-appended entry in that round (reset with the word)
  int *nonapp_list = a.iscr + 6 * (long long)a.p;                     // CTA 0
  double maxH = 0.0;
  if (c.bid == 0 && tid == 0) sm->nonapp = 0;
  // Window of columns evaluated in parallel against the same r.  A mover invalidates everything behind it,
  // so the window restarts small right behind a mover (those columns are L2-hot: a cheap re-evaluation)
  // and doubles after every clean round up to CH (streaming regime: one barrier per CH columns).
  // Right behind a mover the window is G columns, ONE PER CTA (16 warps share a column: lowest latency,
  // movers tend to cluster); a clean round then widens it to one column per warp, 2 per warp, ...
  const int Wwarp = min(c.CH, c.G * NV_W);
  auto grow = [&](const NWin &w) { return w.cta ? Wwarp : min(c.CH, 2 * w.W); };
  auto make = [&](int q0, int W) {
    NWin w;
    w.q0 = q0;
    w.W = W;
    w.qlen = min(W, a.p - q0);
    w.cta = W <= c.G;
    w.slot = (int)(c.rnd & 3u);
    c.rnd += 1;
    return w;
  };
  const bool pipeline = a.pipeline != 0;
  int Wnext = (nact_hint > 0 && mP == 0) ? min(c.CH, c.G) : Wwarp;
  int q0 = q_start, streak = mP > 0 ? 2 : 0; // with a plan the sweep is expected to be clean: pipelined from the first window
  int pi = 0;                          // next planned member
  int seg_events = 0;                  // unplanned movers of this segment
  bool sw = false;                     // `pend` is a planned super-window: entries [pi, sw_pj) applied to r, not yet committed
  int sw_pj = 0;
  double *sw_snap = nullptr, sw_rr = 0.0; // r and ||r||^2 at its start
  q_resume = a.p;
#ifdef CDGPU_WITH_REPLAN
  int replans = 0;                     // plans redone in this pass
#endif
  bool have_pend = false;
  NWin pend = {}, spec = {};
  for (;;) {
    long long ta = clock64();
    if (!have_pend) {
      if (q0 >= a.p) break;
      if (DENSE && dense && truncated && mP > 0 && pi == mP) { // the plan stopped at `cap` entries: plan the rest
        q_resume = q0;
        break;
      }
      // a super-window pays when at least 8 planned entries lie among the next CH positions: every gap then saves a
      // grid barrier; a sparser plan keeps the pipelined windows
      constexpr int sw_min = 8;
      int planned_near = 0;
      if (DENSE)
        for (int t = pi; t < mP && planned_near < sw_min && c.e_row[t] < q0 + c.CH; ++t) planned_near += 1;
      if (pi < mP && planned_near < sw_min && c.e_row[pi] == q0) {
        // sparse plan (a few members in a long streaming pass): a run of planned entries at consecutive positions is
        // one fused update of r between two ordinary windows
        int pj = pi + 1;
        while (pj < mP && c.e_row[pj] == q0 + (pj - pi)) ++pj;
        apply_planned_r(c, pi, pj);
        commit_planned(c, pi, pj, maxH, accepted);
        for (int t = pi; t < pj; ++t) plan_moved += c.e_g[t] != 0.0;
        plan_used += pj - pi;
        q0 += pj - pi;
        pi = pj;
        pf[3] += clock64() - ta;
        continue;
      }
      if (DENSE && pi < mP && planned_near >= sw_min) {
        // PLANNED SUPER-WINDOW (dense plan): up to CH positions, planned and not, in ONE round.  Every CTA walks the window in visit
        // order on its own copy of r: the columns of a gap between planned positions are evaluated (one warp each)
        // against r as it stands there, then the planned run behind the gap is applied — no grid barrier per gap,
        // only two CTA barriers.  The iterate and the list are written once the round is known to be clean
        // (commit_planned); if an unplanned coordinate of the window does move, r is restored from the snapshot CTA 0
        // took at the start (global memory, two buffers) and the planned steps before the mover are replayed: the same
        // operations in the same order, so r is bit for bit what the one-round-per-mover pass has there.
        pend = make(q0, c.CH);
        pend.cta = true; // CTA-wide barrier, no speculative next window
        HEntry *hbw = c.hbuf + (size_t)pend.slot * c.CH;
        sw_snap = a.rsnap + (size_t)(pend.slot & 1) * a.n;
        sw_rr = c.rr;
        if (c.bid == 0)
          for (int i = tid; i < a.n; i += NV_T) __stcg(sw_snap + i, c.r[i]);
        int pj = pi, ja = 0;
        while (ja < pend.qlen) {
          const int jb = (pj < mP && c.e_row[pj] - q0 < pend.qlen) ? c.e_row[pj] - q0 : pend.qlen;
          if (jb > ja) eval_window(c, pend, lam, pk, ordered, words, nonapp_flags, ja, jb, true);
          if (jb < pend.qlen) { // the run of planned entries at consecutive positions starting at jb
            int pk2 = pj + 1;
            while (pk2 < mP && c.e_row[pk2] == c.e_row[pj] + (pk2 - pj) && c.e_row[pk2] - q0 < pend.qlen) ++pk2;
            if (c.bid == 0)
              for (int t = pj + tid; t < pk2; t += NV_T) __stcg(&hbw[c.e_row[t] - q0].app, 1); // (not a window entry)
            __syncthreads(); // every warp has finished reading r for the gap
            apply_planned_r(c, pj, pk2);
            ja = jb + (pk2 - pj);
            pj = pk2;
          } else {
            ja = jb;
          }
        }
        sw_pj = pj;
        sw = true;
        round_arrive_cta(c);
        have_pend = true;
      } else {
        pend = make(q0, Wnext);
        if (pi < mP) pend.qlen = min(pend.qlen, c.e_row[pi] - q0); // windows never cross a planned position
        eval_window(c, pend, lam, pk, ordered, words, nonapp_flags, 0, pend.qlen, false);
        if (pend.cta) round_arrive_cta(c); else round_arrive_warp(c);
        sw = false;
        have_pend = true;
      }
    }
    // the next window, on the assumption that `pend` turns out clean: evaluated while barrier `pend` completes
    bool have_spec = false;
    if (pipeline && streak >= 2 && !pend.cta && pend.q0 + pend.qlen < a.p && !(pi < mP && c.e_row[pi] == pend.q0 + pend.qlen)) {
      spec = make(pend.q0 + pend.qlen, grow(pend));
      if (pi < mP) spec.qlen = min(spec.qlen, c.e_row[pi] - spec.q0);
      eval_window(c, spec, lam, pk, ordered, words, nonapp_flags, 0, spec.qlen, false);
      have_spec = true;
    }
    const long long tb = clock64();
    if (pend.cta) round_wait_cta(c); else round_wait_warp(c);
    const long long tc = clock64();
    pf[0] += tb - ta;
    pf[1] += tc - tb;
    pf[7] += 1;
    const HEntry *hb = c.hbuf + (size_t)pend.slot * c.CH;
    const unsigned int jmin = __ldcg(words + pend.slot);
    if (c.bid == 0 && tid == 0) { // the previous round's word and flag
      __stcg(words + ((pend.slot + 3) & 3), 0xffffffffu);
      __stcg(nonapp_flags + ((pend.slot + 3) & 3), 0);
    }
    if (c.bid == 0 && __ldcg(nonapp_flags + pend.slot)) { // rare: remember the finalised non-appended coordinates
      __syncthreads();
      const int jend = jmin == 0xffffffffu ? pend.qlen : (int)jmin;
      for (int j = tid; j < jend; j += NV_T)
        if (__ldcg(&hb[j].app) == 0)
          nonapp_list[atomicAdd(&sm->nonapp, 1)] = ordered ? pend.q0 + j : (int)cd_perm(pk, (uint32_t)(pend.q0 + j));
      __syncthreads();
    }
    const long long td = clock64();
    pf[2] += td - tc;
    if (jmin == 0xffffffffu) { // clean round: every position of the window is final
      if (DENSE && sw) {
        commit_planned(c, pi, sw_pj, maxH, accepted);
        for (int t = pi; t < sw_pj; ++t) plan_moved += c.e_g[t] != 0.0;
        plan_used += sw_pj - pi;
        pi = sw_pj;
        sw = false;
      }
      q0 = pend.q0 + pend.qlen;
      Wnext = grow(pend);
      streak += 1;
      if (have_spec) {
        round_arrive_warp(c); // arrive_{s+1} only now: the grid counter is shared by consecutive barriers
        pend = spec;
        Wnext = grow(pend);
      } else {
        have_pend = false;
      }
      continue;
    }
    const int k = ordered ? pend.q0 + (int)jmin : (int)cd_perm(pk, (uint32_t)(pend.q0 + jmin));
    const double2 e = __ldcg(reinterpret_cast<const double2 *>(hb + jmin));
    const double h = e.x, nw = e.y;
    if (DENSE && sw) { // an unplanned coordinate of a super-window moved: the planned entries before it stand, the rest is undone
      int pc = pi;
      while (pc < sw_pj && c.e_row[pc] < pend.q0 + (int)jmin) ++pc;
      if (pc < sw_pj) {
        __syncthreads();
        for (int i = tid; i < a.n; i += NV_T) c.r[i] = __ldcg(sw_snap + i);
        c.rr = sw_rr;
        __syncthreads();
        if (pc > pi) apply_planned_r(c, pi, pc);
      }
      commit_planned(c, pi, pc, maxH, accepted);
      for (int t = pi; t < pc; ++t) plan_moved += c.e_g[t] != 0.0;
      plan_used += pc - pi;
      pi = pc;
      sw = false;
    }
    if (!pend.cta) __syncthreads(); // warp-granular round: the other warps may still be reading r
    if (c.bid == 0 && tid == 0) {
      __stcg(a.beta + k, nw);
      if (!a.inlist[k]) { // setindex! appends on the first non-zero store
        __stcg(a.inlist + k, (unsigned char)1);
        a.act[sm->nact] = k;
        sm->nact += 1;
      }
    }
    apply_step(c, a.X + (long long)k * a.ldx, h);
    maxH = fmax(maxH, fabs(h));
    accepted += 1;
    q0 = pend.q0 + (int)jmin + 1;
    Wnext = min(c.CH, c.G);
    streak = 0;
    have_pend = false;
    if (have_spec) { // the discarded round still has its barrier (nothing is read from it)
      round_arrive_cta(c);
      round_wait_cta(c);
      if (c.bid == 0 && tid == 0) {
        __stcg(words + ((spec.slot + 3) & 3), 0xffffffffu);
        __stcg(nonapp_flags + ((spec.slot + 3) & 3), 0);
      }
    }
    pf[3] += clock64() - td;
    // a non-member moved: the steps planned for the members still to come are void.  With enough of them left they are
    // planned again against the new r (fresh d = X_k'(w.r), the same active Gram); otherwise they become ordinary columns.
    // (Compiled in with -DCDGPU_WITH_REPLAN and switched on by CDGPU_NAIVE_REPLAN=1 only: measured on B200 the two grid
    // barriers + chain of a re-plan cost more than the rounds they save — C1 lambda=0.05 8.3 ms with, 6.9 ms without — and
    // the inlined call in this loop costs the streaming rounds of every pass 7 % through register allocation alone.)
#ifdef CDGPU_WITH_REPLAN
    if (mP > 0 && a.replan && mP - pi >= NV_REPLAN_MIN && replans < NV_REPLAN_MAX) {
      const long long tr0 = clock64();
      replan_members(c, lam, pi, mP);
      replans += 1;
      Wnext = Wwarp;
      streak = 2;
      pf[8] += clock64() - tr0;
    } else
#endif
    {
      seg_events += 1;
      events += 1;
      if (DENSE && dense && a.p - q0 >= 64 && (mP - pi >= NV_DENSE_MIN_LEFT || seg_events >= NV_DENSE_EVENTS)) {
        q_resume = q0; // a new segment: plan the rest of the pass against the current r
        break;
      }
      mP = 0;
    }
  }
  __syncthreads(); // warps of a warp-granular round leave together
  return maxH;
}

// dense mode, every CTA: bscr[q] for the visit positions q >= q0 of the pass: 1 = the coordinate is listed, 2 = a
// non-member that would move against the CURRENT r at 0.9 lambda (a candidate), 0 = neither
__device__ void dense_prescan(NCtx &c, double lam, const PermKey &pk, bool ordered, int q0) {
  const NaiveArgs &a = c.a;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int q = q0 + c.bid + c.G * warp; q < a.p; q += c.G * NV_W) {
    const int k = ordered ? q : (int)cd_perm(pk, (uint32_t)q);
    unsigned char f = 1;
    if (!__ldcg(a.inlist + k)) {
      const double d = warp_col_dot(c, a.X + (long long)k * a.ldx);
      double nw, h;
      bool tnz;
      coord_update(a.kind, a.n, d, __ldg(a.colsq + k), __ldcg(a.beta + k), 0.9 * lam, a.omega ? __ldg(a.omega + k) : 1.0, c.rr,
                   nw, h, tnz);
      f = h != 0.0 ? 2 : 0;
    }
    if (lane == 0) __stcg(a.bscr + q, f);
  }
}

// dense mode, CTA 0: the flagged positions from q0 on, in visit order, become the plan's coordinate list (at most `cap`)
__device__ void dense_select(NCtx &c, const PermKey &pk, bool ordered, int q0, int cap) {
  const NaiveArgs &a = c.a;
  NSmem *sm = c.sm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const NPlan PL = plan_arrays(a);
  int cnt = 0;
  for (int base = q0; base < a.p && cnt <= cap; base += NV_T) {
    const int q = base + tid;
    const bool f = q < a.p && __ldcg(a.bscr + q) != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    if (lane == 0) sm->redu[warp] = __popc(bal);
    __syncthreads();
    int off = 0, tot = 0;
    for (int w2 = 0; w2 < NV_W; ++w2) {
      const int v = (int)sm->redu[w2];
      if (w2 < warp) off += v;
      tot += v;
    }
    const int pos = cnt + off + __popc(bal & ((1u << lane) - 1u));
    if (f && pos < cap) PL.ext[pos] = ordered ? q : (int)cd_perm(pk, (uint32_t)q);
    cnt += tot;
    __syncthreads();
  }
  if (tid == 0) {
    PL.row[513] = min(cnt, cap);
    PL.row[514] = cnt > cap ? 1 : 0;
  }
  __threadfence();
  __syncthreads();
}

// dropzeros! after a full pass on CTA 0: new entries go where the reference's temporary appends put
// them (common.cuh: cd_compact_list); sqrt-lasso never appends temporarily (:278-283).
__device__ void list_update_full(NCtx &c, int m_old, unsigned long long pass_counter) {
  const NaiveArgs &a = c.a;
  NSmem *sm = c.sm;
  const int tid = threadIdx.x;
  const int m = sm->nact;
  for (int i = tid; i < m; i += NV_T) a.actval[i] = __ldcg(a.beta + a.act[i]);
  int *newpos = a.iscr + 7 * (long long)a.p;
  const int *nonapp = a.iscr + 6 * (long long)a.p;
  const int nna = sm->nonapp;
  const bool ordered = a.randomize == 0;
  const PermKey pk = cd_perm_key((uint32_t)a.p, a.seed, pass_counter);
  for (int e = m_old + tid; e < m; e += NV_T) {
    if (a.kind == CDGPU_LOSS_SQRT) {
      newpos[e - m_old] = e;
      continue;
    }
    const int k = a.act[e];
    const long long vis = ordered ? k : (long long)cd_perm_inv(pk, (uint32_t)k);
    int before = 0;
    for (int j = 0; j < m_old; ++j) {
      const int kj = a.act[j];
      before += (ordered ? kj : (long long)cd_perm_inv(pk, (uint32_t)kj)) < vis;
    }
    for (int j = 0; j < nna; ++j) {
      const int kj = nonapp[j];
      before += (ordered ? kj : (long long)cd_perm_inv(pk, (uint32_t)kj)) < vis;
    }
    newpos[e - m_old] = m_old + (int)vis - before;
  }
  __syncthreads();
  cd_compact_list<NV_T>(a.act, a.actval, m_old, m, newpos, a.inlist, a.iscr + (long long)a.p, a.scr + 8 + 8 * (long long)a.p + 16,
                        sm->s2);
  if (tid == 0) sm->nact = sm->s2[0];
  __syncthreads();
}

// consecutive active-set passes on CTA 0 (sequential chain; one column per step)
__device__ void active_phase(NCtx &c, double lam, long long maxPasses, unsigned long long pass_counter) {
  const NaiveArgs &a = c.a;
  NSmem *sm = c.sm;
  const int tid = threadIdx.x, n = a.n;
  const bool ordered = a.randomize == 0;
  long long npasses = 0, visits = 0, accepted = 0;
  double maxH = 0.0;
  int conv = 0;
  while (npasses < maxPasses) {
    const int m = sm->nact;
    const PermKey pkm = cd_perm_key((uint32_t)max(m, 1), a.seed, pass_counter + npasses);
    double pmax = 0.0;
    for (int s = 0; s < m; ++s) {
      const int i = ordered ? s : (int)cd_perm(pkm, (uint32_t)s);
      const int k = a.act[i];
      const double *col = a.X + (long long)k * a.ldx;
      if (s + 1 < m) { // pull the next column towards the SM while this one is reduced
        const int kn = a.act[ordered ? s + 1 : (int)cd_perm(pkm, (uint32_t)(s + 1))];
        const char *nc = reinterpret_cast<const char *>(a.X + (long long)kn * a.ldx);
        for (int off = tid * 128; off < n * 8; off += NV_T * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nc + off));
      }
      double d = 0.0;
      for (int t0 = tid; t0 < n; t0 += 8 * NV_T) { // 8 independent loads in flight per thread
        double xv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) xv[u] = (t0 + u * NV_T < n) ? __ldg(col + t0 + u * NV_T) : 0.0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int t = t0 + u * NV_T;
          if (t < n) d = fma(c.w ? xv[u] * c.w[t] : xv[u], c.r[t], d);
        }
      }
      d = block_sum(sm, d, 1);
      double nw, h;
      bool tnz;
      const double old = a.actval[i];
      coord_update(a.kind, n, d, __ldg(a.colsq + k), old, lam, a.omega ? __ldg(a.omega + k) : 1.0, c.rr, nw, h, tnz);
      __syncthreads(); // everyone has read actval[i] and red[1]
      if (tid == 0) a.actval[i] = nw;
      if (h != 0.0) {
        apply_step(c, col, h);
        accepted += 1;
      }
      pmax = fmax(pmax, fabs(h));
    }
    npasses += 1;
    visits += m;
    maxH = pmax;
    // dropzeros!
    __syncthreads();
    int anyz = 0;
    for (int i = tid; i < m; i += NV_T) anyz |= (a.actval[i] == 0.0);
    anyz = __syncthreads_or(anyz);
    if (anyz) {
      if (tid == 0) {
        int nn = m, i = 0;
        while (i < nn) {
          if (a.actval[i] == 0.0) {
            a.inlist[a.act[i]] = 0;
            __stcg(a.beta + a.act[i], 0.0);
            if (i != nn - 1) {
              a.actval[i] = a.actval[nn - 1];
              a.act[i] = a.act[nn - 1];
            }
            nn -= 1;
          } else {
            i += 1;
          }
        }
        sm->nact = nn;
      }
      __syncthreads();
    }
    if (maxH < a.optTol) {
      conv = 1;
      break;
    }
  }
  // publish: dense beta of the listed coordinates, residual, summary
  const int m = sm->nact;
  for (int i = tid; i < m; i += NV_T) __stcg(a.beta + a.act[i], a.actval[i]);
  for (int i = tid; i < n; i += NV_T) __stcg(a.r + i, c.r[i]);
  if (tid == 0) {
    c.bc->npasses = npasses;
    c.bc->visits = visits;
    c.bc->accepted = accepted;
    c.bc->maxH = maxH;
    c.bc->conv = conv;
    c.bc->nact = m;
  }
  __threadfence();
  __syncthreads();
}

// ------------------------------------------------- active-set passes in covariance form --
// An active-set pass only touches the m active columns, and X_k'(w.r) for k in the set obeys
// d_t -= G_tk h with G = X_A' diag(w) X_A.  So the whole grid forms G (m x m) and d = X_A'(w.r) once per
// active phase (m(m+1)/2 + m dot products, one warp each, columns are L2 hot), CTA 0 then runs the
// sequential chain with one register-resident entry per thread, one named barrier per coordinate
// step and the G column gathers prefetched two steps ahead (no O(n) work per step at all), and
// finally r -= X_A (beta - beta_at_entry) is applied once.  Same iterates as the reference up to
// rounding (d maintained incrementally instead of re-reduced); same visit order and list semantics.
constexpr int NV_GCAP = NV_GCAP_; // largest active set handled this way (G scratch = gram_cap^2 doubles: 134 MB at 4096)

__device__ __forceinline__ void nbar(int id, int nthr) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthr) : "memory"); }

// all CTAs: G[i + j*m] = sum_t w_t X[t,act_i] X[t,act_j] (both triangles), d[i] = sum_t w_t r_t X[t,act_i]
__device__ void build_active_gram(NCtx &c, int m, double *G, double *d, const int *lst) {
  const NaiveArgs &a = c.a;
  const int lane = threadIdx.x & 31, n = a.n;
  const long long gw = (long long)c.bid * NV_W + (threadIdx.x >> 5), nw = (long long)c.G * NV_W;
  const long long npair = (long long)m * (m + 1) / 2;
  for (long long idx = gw; idx < npair + m; idx += nw) {
    if (idx >= npair) { // d entry
      const int i = (int)(idx - npair);
      const double v = warp_col_dot(c, a.X + (long long)__ldcg(lst + i) * a.ldx);
      if (lane == 0) __stcg(d + i, v);
      continue;
    }
    // idx -> (i, j), j <= i, row-major over the lower triangle
    int i = (int)((sqrt(8.0 * (double)idx + 1.0) - 1.0) * 0.5);
    while ((long long)i * (i + 1) / 2 > idx) --i;
    while ((long long)(i + 1) * (i + 2) / 2 <= idx) ++i;
    const int j = (int)(idx - (long long)i * (i + 1) / 2);
    const double *ci = a.X + (long long)__ldcg(lst + i) * a.ldx, *cj = a.X + (long long)__ldcg(lst + j) * a.ldx;
    double s0 = 0.0, s1 = 0.0;
    if (c.w) {
      int t = lane;
      for (; t + 32 < n; t += 64) {
        s0 = fma(__ldg(ci + t) * c.w[t], __ldg(cj + t), s0);
        s1 = fma(__ldg(ci + t + 32) * c.w[t + 32], __ldg(cj + t + 32), s1);
      }
      for (; t < n; t += 32) s0 = fma(__ldg(ci + t) * c.w[t], __ldg(cj + t), s0);
    } else {
      int t = lane;
      for (; t + 32 < n; t += 64) {
        s0 = fma(__ldg(ci + t), __ldg(cj + t), s0);
        s1 = fma(__ldg(ci + t + 32), __ldg(cj + t + 32), s1);
      }
      for (; t < n; t += 32) s0 = fma(__ldg(ci + t), __ldg(cj + t), s0);
    }
    const double v = warp_sum(s0 + s1);
    if (lane == 0) {
      __stcg(G + i + (long long)j * m, v);
      __stcg(G + j + (long long)i * m, v);
    }
  }
}

// CTA 0 after an active phase on the Gram: publish the new iterate (list, dense beta, membership) and fold the change
// into the residual once, r -= X[:, act0] (beta - beta_at_entry); summary to the broadcast block.
__device__ void gram_publish(NCtx &c, const chain::Result &r, int m0, const int *act0, const double *scr_b0, double *scr_dlt) {
  const NaiveArgs &a = c.a;
  NSmem *sm = c.sm;
  const int tid = threadIdx.x, n = a.n;
  const int m = r.m;
  for (int i = tid; i < m0; i += NV_T) __stcg(a.beta + act0[i], 0.0);
  __syncthreads();
  for (int i = tid; i < m; i += NV_T) {
    const int k = c.e_coord[i];
    const double be = c.e_be[i];
    __stcg(a.beta + k, be);
    a.act[i] = k;
    a.actval[i] = be;
    __stcg(a.inlist + k, (unsigned char)1);
  }
  __syncthreads();
  for (int i = tid; i < m0; i += NV_T) scr_dlt[i] = __ldcg(a.beta + act0[i]) - scr_b0[i];
  __syncthreads();
  for (int t0 = tid; t0 < n; t0 += NV_T) {
    const double *row = a.X + t0;
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
    int i = 0;
    for (; i + 4 <= m0; i += 4) {
      const double v0 = __ldg(row + (long long)act0[i] * a.ldx), v1 = __ldg(row + (long long)act0[i + 1] * a.ldx);
      const double v2 = __ldg(row + (long long)act0[i + 2] * a.ldx), v3 = __ldg(row + (long long)act0[i + 3] * a.ldx);
      acc0 = fma(v0, scr_dlt[i], acc0);
      acc1 = fma(v1, scr_dlt[i + 1], acc1);
      acc2 = fma(v2, scr_dlt[i + 2], acc2);
      acc3 = fma(v3, scr_dlt[i + 3], acc3);
    }
    for (; i < m0; ++i) acc0 = fma(__ldg(row + (long long)act0[i] * a.ldx), scr_dlt[i], acc0);
    const double v = c.r[t0] - ((acc0 + acc1) + (acc2 + acc3));
    c.r[t0] = v;
    __stcg(a.r + t0, v);
  }
  if (tid == 0) {
    sm->nact = m;
    c.bc->npasses = r.npasses;
    c.bc->visits = r.visits;
    c.bc->accepted = r.accepted;
    c.bc->maxH = r.maxH;
    c.bc->conv = r.conv;
    c.bc->nact = m;
  }
  __threadfence();
  __syncthreads();
}

// closed-form updates on d_t = X_t'(w.r) for the chain engine (chain_engine.cuh)
struct LsPolicy { // cd_differentiable_function.jl:101-104 / :184-187
  static constexpr bool HAS_RR = false;
  static constexpr bool FAST_V = true; // chain_engine.cuh: chain_steps
  const double *colsq, *omega;
  double lam, nd;
  __device__ __forceinline__ void load_consts(int k, double &c0, double &c1, double &c2) const {
    c0 = __ldg(colsq + k);
    c1 = 1.0 / c0;
    c2 = __dmul_rn(__dmul_rn(nd / c0, lam), omega ? __ldg(omega + k) : 1.0);
  }
  // d / aa with the reciprocal known: one Newton correction of the product gives the correctly rounded quotient (the
  // sequence a division expands to, minus the reciprocal: ~25 instead of ~120 cycles on the dependent chain); operands
  // outside its safe range (tiny / huge quotients, non-finite values, a zero column) take the plain division
  static __device__ __forceinline__ double quot(double d, double aa, double inv) {
    const double q0 = d * inv;
    const double e = fma(-q0, aa, d);
    const double q1 = fma(e, inv, q0);
    const double aq = fabs(q1);
    return (aq < 1e280 && (aq > 1e-280 || d == 0.0)) ? q1 : d / aa;
  }
  // argument of the shrinkage, as step() forms it
  __device__ __forceinline__ double enter(double d, double be, double aa, double inv) const { return __dadd_rn(be, quot(d, aa, inv)); }
  __device__ __forceinline__ void step(double d, double be, double aa, double inv, double th, double, double &nw, double &h,
                                       double &dr) const {
    const double v = __dadd_rn(be, quot(d, aa, inv));
    nw = cd_shrink(v, th);
    h = nw - be;
    dr = 0.0;
  }
  static __device__ __forceinline__ double apply(double d, double Gv, double h) { return __dsub_rn(d, __dmul_rn(Gv, h)); }
};
struct SqrtPolicy { // :259-283 with s, ||r+||^2 from d, ||r||^2, a
  static constexpr bool HAS_RR = true;
  static constexpr bool FAST_V = false; // the threshold moves with ||r||^2: every step is evaluated in full
  __device__ __forceinline__ double enter(double, double, double, double) const { return 0.0; }
  const double *colsq, *omega;
  double lam;
  __device__ __forceinline__ void load_consts(int k, double &c0, double &c1, double &c2) const {
    c0 = __ldg(colsq + k);
    c2 = lam * (omega ? __ldg(omega + k) : 1.0);
    c1 = c2 / sqrt(1.0 - c2 * c2 / c0); // the state-independent factor of :280-283, same operation order
  }
  __device__ __forceinline__ void step(double d, double old, double aa, double lq, double l, double rr, double &nw, double &h,
                                       double &dr) const {
    const double sv = d + aa * old;
    const double rsq = rr + 2.0 * old * d + old * old * aa;
    const double t = l * sqrt(rsq);
    if (fabs(sv) <= t)
      nw = 0.0;
    else if (sv > t)
      nw = (sv - lq * sqrt(rsq - sv * sv / aa)) / aa;
    else
      nw = (sv + lq * sqrt(rsq - sv * sv / aa)) / aa;
    h = nw - old;
    dr = h * (h * aa - 2.0 * d); // change of ||r||^2
  }
  static __device__ __forceinline__ double apply(double d, double Gv, double h) { return __dsub_rn(d, __dmul_rn(Gv, h)); }
};

// CTA 0, before a full pass: the steps the m members will take when the pass reaches them, by ONE chain pass over the
// active Gram G = X_A'[W]X_A and d = X_A'(w.r) (just formed by the grid) in the VISIT order of the full pass, on the
// assumption that no non-member moves (full_pass drops the rest of the plan when one does).  Plan -> global memory.
__device__ void member_plan(NCtx &c, double lam, unsigned long long pass_counter, int m, const double *G, const double *d0,
                            const int *lst) {
  const NaiveArgs &a = c.a;
  NSmem *sm = c.sm;
  const int tid = threadIdx.x;
  const bool ordered = a.randomize == 0;
  const PermKey pk = cd_perm_key((uint32_t)a.p, a.seed, pass_counter);
  const NPlan PL = plan_arrays(a);
  int *vis = c.e_coord; // visit position by list index (temporary)
  for (int i = tid; i < m; i += NV_T) {
    const int k = __ldcg(lst + i);
    vis[i] = ordered ? k : (int)cd_perm_inv(pk, (uint32_t)k);
  }
  __syncthreads();
  for (int i = tid; i < m; i += NV_T) {
    const int my = vis[i];
    int t = 0;
    for (int j = 0; j < m; ++j) t += vis[j] < my;
    c.e_row[t] = i; // t-th visited member = list entry i = row i of G
    PL.pos[t] = my;
    PL.row[t] = i;
  }
  __syncthreads();
  for (int t = tid; t < m; t += NV_T) {
    const int i = c.e_row[t];
    c.e_be[t] = __ldcg(a.beta + __ldcg(lst + i)); // (the dense iterate is current; candidates: 0)
    c.e_g[t] = __ldcg(d0 + i);
  }
  __syncthreads();
  for (int t = tid; t < m; t += NV_T) {
    const int k = __ldcg(lst + c.e_row[t]);
    c.e_coord[t] = k;
    PL.k[t] = k;
  }
  __syncthreads();
  chain::State S;
  S.m = m;
  S.row = c.e_row;
  S.coord = c.e_coord;
  S.g = c.e_g;
  S.be = c.e_be;
  S.ord = c.e_ord;
  S.pos = c.e_pos;
  S.stage = c.e_stage;
  S.sh = &sm->ch;
  S.G = G;
  S.ldg = m;
  S.prof = nullptr;
  S.hout = PL.h;
  if (a.kind == CDGPU_LOSS_SQRT) {
    const SqrtPolicy P{a.colsq, a.omega, lam};
    (void)chain::run<NV_T>(S, P, c.rr, 1, pass_counter, true, a.seed, a.optTol, nullptr);
  } else {
    const LsPolicy P{a.colsq, a.omega, lam, (double)a.n};
    (void)chain::run<NV_T>(S, P, 0.0, 1, pass_counter, true, a.seed, a.optTol, nullptr);
  }
  __syncthreads();
  for (int t = tid; t < m; t += NV_T) PL.nw[t] = c.e_be[t];
  if (tid == 0) PL.row[512] = m; // leading dimension of G, for replan_members
  __threadfence();
  __syncthreads();
}

// Every CTA, in the middle of a planned full pass, after a non-member has moved: the steps of the planned members
// [t0, t1) still to come are computed again.  d = X_k'(w.r) fresh against the current r (one warp per member, all CTAs),
// then CTA 0 runs the chain over them in visit order on the active Gram formed for this pass (the new coordinate has
// been visited already and takes no further part), and everybody reloads that part of the plan.
__device__ void replan_members(NCtx &c, double lam, int t0, int t1) {
  const NaiveArgs &a = c.a;
  NSmem *sm = c.sm;
  const int tid = threadIdx.x, lane = tid & 31;
  const NPlan PL = plan_arrays(a);
  const double *G = a.gram;
  double *ds = a.gram + (long long)a.gram_cap * a.gram_cap;
  const int mlist = (int)__ldcg(reinterpret_cast<const int *>(PL.row + 512)); // list length the Gram was formed for
  {
    const int gw = c.bid * NV_W + (tid >> 5), nw = c.G * NV_W;
    for (int t = t0 + gw; t < t1; t += nw) {
      const double v = warp_col_dot(c, a.X + (long long)c.e_coord[t] * a.ldx);
      if (lane == 0) __stcg(ds + t, v);
    }
  }
  fast_grid_sync(c);
  if (c.bid == 0) {
    const int m = t1 - t0; // <= NV_PLAN_MAX <= NV_T: one entry per thread
    int k = 0, row = 0;
    double be = 0.0, g = 0.0;
    if (tid < m) {
      k = c.e_coord[t0 + tid];
      row = __ldcg(PL.row + t0 + tid);
      be = a.actval[row]; // the members still to come have not changed in this pass
      g = __ldcg(ds + t0 + tid);
    }
    __syncthreads();
    if (tid < m) {
      c.e_coord[tid] = k;
      c.e_row[tid] = row;
      c.e_be[tid] = be;
      c.e_g[tid] = g;
    }
    __syncthreads();
    chain::State S;
    S.m = m;
    S.row = c.e_row;
    S.coord = c.e_coord;
    S.g = c.e_g;
    S.be = c.e_be;
    S.ord = c.e_ord;
    S.pos = c.e_pos;
    S.stage = c.e_stage;
    S.sh = &sm->ch;
    S.G = G;
    S.ldg = mlist;
    S.prof = nullptr;
    S.hout = PL.h + t0;
    if (a.kind == CDGPU_LOSS_SQRT) {
      const SqrtPolicy P{a.colsq, a.omega, lam};
      (void)chain::run<NV_T>(S, P, c.rr, 1, 0, true, a.seed, a.optTol, nullptr);
    } else {
      const LsPolicy P{a.colsq, a.omega, lam, (double)a.n};
      (void)chain::run<NV_T>(S, P, 0.0, 1, 0, true, a.seed, a.optTol, nullptr);
    }
    __syncthreads();
    if (tid < m) PL.nw[t0 + tid] = c.e_be[tid];
    __threadfence();
    __syncthreads();
  }
  fast_grid_sync(c);
  for (int t = t0 + tid; t < t1; t += NV_T) {
    c.e_coord[t] = __ldcg(PL.k + t);
    c.e_row[t] = __ldcg(PL.pos + t);
    c.e_g[t] = __ldcg(PL.h + t);
    c.e_be[t] = __ldcg(PL.nw + t);
  }
  __syncthreads();
}

// CTA 0: the active-set passes as the blocked warp-level chain over (d, beta) of the m0 stored entries
__device__ void gram_engine(NCtx &c, double lam, long long maxPasses, unsigned long long pass_counter, int m0,
                            const double *G, const double *d0) {
  const NaiveArgs &a = c.a;
  NSmem *sm = c.sm;
  const int tid = threadIdx.x, n = a.n;
  double *scr_b0 = a.scr + 8 + 9 * (long long)a.p + 32; // [<= NV_GCAP] each, behind the compaction staging
  double *scr_dlt = scr_b0 + NV_GCAP;
  int *act0 = a.iscr; // snapshot of the list
  for (int i = tid; i < m0; i += NV_T) {
    const int k = a.act[i];
    const double be = a.actval[i];
    c.e_coord[i] = k;
    c.e_row[i] = i;
    act0[i] = k;
    c.e_be[i] = be;
    scr_b0[i] = be;
    c.e_g[i] = __ldcg(d0 + i);
  }
  __syncthreads();
  chain::State S;
  S.m = m0;
  S.row = c.e_row;
  S.coord = c.e_coord;
  S.g = c.e_g;
  S.be = c.e_be;
  S.ord = c.e_ord;
  S.pos = c.e_pos;
  S.stage = c.e_stage;
  S.sh = &sm->ch;
  S.G = G;
  S.ldg = m0;
  S.prof = a.prof ? a.prof + 16 : nullptr;
  chain::Result r;
  const bool ordered = a.randomize == 0;
  if (a.kind == CDGPU_LOSS_SQRT) {
    const SqrtPolicy P{a.colsq, a.omega, lam};
    r = chain::run<NV_T>(S, P, c.rr, maxPasses, pass_counter, ordered, a.seed, a.optTol, a.inlist);
  } else {
    const LsPolicy P{a.colsq, a.omega, lam, (double)n};
    r = chain::run<NV_T>(S, P, 0.0, maxPasses, pass_counter, ordered, a.seed, a.optTol, a.inlist);
  }
  gram_publish(c, r, m0, act0, scr_b0, scr_dlt);
}

// The same phase with the chain engine spread over the whole cooperative grid (chain_engine.cuh: run_multi), for
// active sets of many 32-entry blocks; called by ALL CTAs after the active Gram has been formed.
constexpr int NV_MULTI_TEAM = 16; // CTAs that share the engine
__device__ void gram_engine_multi(NCtx &c, double lam, long long maxPasses, unsigned long long pass_counter, int m0,
                                  const double *G, double *d0) {
  const NaiveArgs &a = c.a;
  NSmem *sm = c.sm;
  const int tid = threadIdx.x, n = a.n;
  double *scr_b0 = a.scr + 8 + 9 * (long long)a.p + 32;
  double *scr_dlt = scr_b0 + NV_GCAP;
  double *hG = scr_b0 + 4 * NV_GCAP + 8, *pmaxG = hG + 64;
  int *flagsG = reinterpret_cast<int *>(hG + 72);
  int *act0 = a.iscr, *rowG = a.iscr + a.p;
  for (int i = tid; i < m0; i += NV_T) c.e_row[i] = i; // every CTA: row ids of the compact Gram
  if (c.bid == 0) {
    for (int i = tid; i < m0; i += NV_T) {
      const int k = a.act[i];
      const double be = a.actval[i];
      c.e_coord[i] = k;
      act0[i] = k;
      c.e_be[i] = be;
      scr_b0[i] = be;
      rowG[i] = i;
    }
  }
  __syncthreads();
  chain::State S;
  S.m = m0;
  S.row = c.e_row;
  S.coord = c.e_coord;
  S.g = nullptr;
  S.be = c.e_be;
  S.ord = c.e_ord;
  S.pos = c.e_pos;
  S.stage = c.e_stage;
  S.sh = &sm->ch;
  S.G = G;
  S.ldg = m0;
  S.prof = a.prof ? a.prof + 16 : nullptr;
  // team = the first W CTAs of the cooperative grid (all co-resident), with its own barrier: one atomic arrive and
  // an acquire-poll on a counter in global memory — a 16-CTA barrier costs a fraction of grid.sync() over 148
  const int W = min(c.G, NV_MULTI_TEAM);
  chain::Multi X{W, c.bid, d0, hG, pmaxG, flagsG, rowG, reinterpret_cast<uint4 *>(a.chain_scr), reinterpret_cast<uint4 *>(a.chain_scr) + CD_GCAP}; // g = X_A'(w.r), already in global memory
  unsigned *ctr = reinterpret_cast<unsigned *>(flagsG + 2);
  unsigned target = 0;
  auto sync = [ctr, &target, W]() {
    __syncthreads();
    if (threadIdx.x == 0) {
      target += (unsigned)W;
      __threadfence();
      atomicAdd(ctr, 1u);
      unsigned v;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      } while (v < target);
    }
    __syncthreads();
  };
  chain::Result r;
  const bool ordered = a.randomize == 0;
  if (a.kind == CDGPU_LOSS_SQRT) {
    const SqrtPolicy P{a.colsq, a.omega, lam};
    r = chain::run_multi<NV_T>(S, X, P, sync, c.rr, maxPasses, pass_counter, ordered, a.seed, a.optTol, a.inlist);
  } else {
    const LsPolicy P{a.colsq, a.omega, lam, (double)n};
    r = chain::run_multi<NV_T>(S, X, P, sync, 0.0, maxPasses, pass_counter, ordered, a.seed, a.optTol, a.inlist);
  }
  if (c.bid != 0) return;
  const long long tpub = clock64();
  gram_publish(c, r, m0, act0, scr_b0, scr_dlt);
  if (a.prof && tid == 0) a.prof[11] += clock64() - tpub;
}

__device__ double shared_std(NCtx &c) { // Statistics.std(r), corrected, two-pass
  const int n = c.a.n;
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += NV_T) s += c.r[i];
  const double mean = block_sum(c.sm, s, 0) / (double)n;
  __syncthreads();
  double v = 0.0;
  for (int i = threadIdx.x; i < n; i += NV_T) v = fma(c.r[i] - mean, c.r[i] - mean, v);
  v = block_sum(c.sm, v, 0);
  __syncthreads();
  return sqrt(v / (double)(n - 1));
}
__device__ double shared_sumsq(NCtx &c) {
  double v = 0.0;
  for (int i = threadIdx.x; i < c.a.n; i += NV_T) v = fma(c.r[i], c.r[i], v);
  v = block_sum(c.sm, v, 0);
  __syncthreads();
  return v;
}

template <bool DENSE>
__global__ void __launch_bounds__(NV_T, 1) naive_path_kernel(const NaiveArgs a, int CH, HEntry *hbuf, NBcast *bc, int gcap) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  NCtx c{a, grid};
  c.sm = reinterpret_cast<NSmem *>(smem_raw);
  unsigned char *sp = smem_raw + (sizeof(NSmem) + 15) / 16 * 16;
  c.gcap = gcap;
  c.e_stage = reinterpret_cast<double *>(sp);
  sp += gcap ? chain::STAGE_DOUBLES * sizeof(double) : 0;
  c.e_g = reinterpret_cast<double *>(sp);
  sp += (size_t)gcap * sizeof(double);
  c.e_be = reinterpret_cast<double *>(sp);
  sp += (size_t)gcap * sizeof(double);
  c.e_row = reinterpret_cast<int *>(sp);
  sp += (size_t)gcap * sizeof(int);
  c.e_coord = reinterpret_cast<int *>(sp);
  sp += (size_t)gcap * sizeof(int);
  c.e_ord = reinterpret_cast<unsigned short *>(sp);
  sp += (size_t)gcap * sizeof(unsigned short);
  c.e_pos = reinterpret_cast<unsigned short *>(sp);
  sp += (size_t)gcap * sizeof(unsigned short);
  double *sd = reinterpret_cast<double *>(sp);
  c.r = sd;
  c.w = a.w ? sd + ((a.n + 1) & ~1) : nullptr; // keep w 16-byte aligned
  c.hbuf = hbuf;
  c.bc = bc;
  c.CH = CH;
  c.G = gridDim.x;
  c.bid = blockIdx.x;
  const int tid = threadIdx.x, n = a.n;
  for (int i = tid; i < n; i += NV_T) {
    c.r[i] = a.r[i];
    if (c.w) c.w[i] = a.w[i];
  }
  if (tid == 0) c.sm->nact = *a.nact;
  c.bar_target = 0;
  c.rnd = 0;
  if (tid == 0) c.sm->warr = 0;
  if (c.bid == 0 && tid == 0) __stcg(reinterpret_cast<unsigned *>(a.flag + 7), 0u);
  __syncthreads();
  grid.sync(); // the round-barrier counter is zero before anyone arrives
  c.rr = shared_sumsq(c);

  long long pf[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (a.prof && c.bid == 0 && tid < 22) a.prof[10 + tid] = 0;
  const long long t_start = clock64();
  int nact_hint = *a.nact; // every CTA's view of the list length (refreshed whenever CTA 0 publishes it)
  int last_events = 0;     // unplanned movers of the last full pass (dense mode of the next one)
  bool hint_stale = false; // a full pass has run since the last refresh
  unsigned long long pass_counter = 0;
  DevStats st;
  st.passes = st.full_passes = st.visits = st.accepted = 0;
  st.maxH = 0.0;
  st.converged = 0;
  st.outer_iters = 0;
  st.sigma = 0.0;
  int status = 0;
  long long cols_done = 0, out_off = 0;
  double sigma = 1.0;
  if (a.scaled == 1) sigma = a.sigma0;
  if (a.scaled == 2) sigma = shared_std(c); // :WarmStart  lasso.jl:124-126
  const long long nouter = a.scaled ? a.outerMaxIter : 1;
  for (int li = 0; li < a.nlambda && status == 0; ++li) {
    if (li > 0 && !a.accumulate) {
      st.passes = st.full_passes = st.visits = st.accepted = 0;
      st.maxH = 0.0;
    }
    for (long long outer = 1; outer <= nouter; ++outer) {
      const double lam = a.scaled ? a.lambdas[li] * sigma : a.lambdas[li];
      st.converged = 0;
      bool conv = true;
      long long iter = 0;
      while (iter < a.maxIter) {
        if (conv) {
          iter += 1;
          st.passes += 1;
          st.full_passes += 1;
          st.visits += a.p;
          const int m_old = c.sm->nact; // CTA 0
          // dense mode (full_pass): when a pre-scan of the remaining columns is cheap (X stays in L2) or the last full
          // pass had many unplanned movers
          if (hint_stale && a.plan && a.gram && c.gcap > 0) { // the list may have changed in the full pass that ended the previous solve
            fast_grid_sync(c);
            nact_hint = __ldcg(&bc->nact);
            hint_stale = false;
          }
          // DENSE instance (X in L2: a pre-scan costs about one round): an empty list (cold start: nothing else can be
          // planned) or 8 unplanned movers in the last full pass
          // (lists longer than one plan are left to the ordinary pass: a 512 x 512 Gram per chunk costs what it saves)
          const bool dense = DENSE && a.dense && a.plan && a.gram && c.gcap > 0 && nact_hint <= min(c.gcap, NV_PLAN_MAX) &&
                             (nact_hint == 0 || last_events >= 8);
          const PermKey pk_pass = cd_perm_key((uint32_t)a.p, a.seed, pass_counter);
          double maxH = 0.0;
          int q_start = 0, events_pass = 0;
          // entries per plan: a cold start begins small (against r = y almost every coordinate looks like a mover, and the
          // first large steps change that), then the size follows how much of the last plan actually moved
          int cap = nact_hint == 0 ? 32 : min(c.gcap, NV_PLAN_MAX);
          int segs = 0;
          for (;;) { // segments of the pass (one, unless dense mode plans again)
            int mP = 0;
            bool truncated = false;
            // at most 6 planned segments per pass (32 + 64 + ... + 512 entries cover a thousand movers): a pass with more
            // movers than that (C1 at lambda = 0.01: a cold pass with 2000) is cheaper one round per mover from there on
            const bool dense_now = dense && segs < 6;
            segs += 1;
            if (a.plan && a.gram && c.gcap > 0) {
              const long long tp0 = clock64();
              double *Gs = a.gram, *ds = a.gram + (long long)a.gram_cap * a.gram_cap;
              const NPlan PL = plan_arrays(a);
              const int *lst = nullptr;
              int mE = 0;
              if (DENSE && dense_now) {
                dense_prescan(c, lam, pk_pass, a.randomize == 0, q_start);
                fast_grid_sync(c);
                if (c.bid == 0) dense_select(c, pk_pass, a.randomize == 0, q_start, cap);
                fast_grid_sync(c);
                mE = __ldcg(PL.row + 513);
                truncated = __ldcg(PL.row + 514) != 0;
                lst = PL.ext;
              } else if (q_start == 0 && nact_hint >= 1 && nact_hint <= min(c.gcap, NV_PLAN_MAX)) {
                mE = nact_hint;
                lst = a.act;
              }
              if (mE >= 1) {
                build_active_gram(c, mE, Gs, ds, lst);
                fast_grid_sync(c);
                if (c.bid == 0) member_plan(c, lam, pass_counter, mE, Gs, ds, lst);
                fast_grid_sync(c);
                mP = mE;
                for (int t = tid; t < mP; t += NV_T) {
                  c.e_coord[t] = __ldcg(PL.k + t);
                  c.e_row[t] = __ldcg(PL.pos + t);
                  c.e_g[t] = __ldcg(PL.h + t);
                  c.e_be[t] = __ldcg(PL.nw + t);
                }
                __syncthreads();
              }
              pf[8] += clock64() - tp0;
            }
            int q_resume = a.p, used = 0, moved = 0;
            maxH = fmax(maxH, full_pass<DENSE>(c, lam, pass_counter, st.accepted, pf, nact_hint, mP, q_start, dense_now, truncated, q_resume,
                                        events_pass, used, moved));
            if (q_resume >= a.p) break;
            q_start = q_resume;
            if (used > 0) cap = 2 * moved >= used ? min(min(c.gcap, NV_PLAN_MAX), 2 * cap) : (4 * moved < used ? max(32, cap / 2) : cap);
          }
          last_events = events_pass;
          hint_stale = true;
          const long long t1 = clock64();
          if (c.bid == 0) {
            list_update_full(c, m_old, pass_counter);
            if (tid == 0) {
              __stcg(&bc->nact, c.sm->nact);
              __threadfence();
            }
          }
          pf[4] += clock64() - t1;
          pass_counter += 1;
          st.maxH = maxH;
          conv = maxH < a.optTol;
          if (conv) {
            st.converged = 1;
            break;
          }
        } else {
          const long long t0 = clock64();
          fast_grid_sync(c); // CTA 0 has published the list length
          const int m_act = __ldcg(&bc->nact);
          nact_hint = m_act;
          hint_stale = false;
          if (m_act >= 1 && m_act <= c.gcap && a.gram) {
            double *Gs = a.gram, *ds = a.gram + (long long)a.gram_cap * a.gram_cap;
            const long long tg0 = clock64();
            build_active_gram(c, m_act, Gs, ds, a.act);
            pf[9] += clock64() - tg0;
            if (c.bid == 0 && tid == 0) { // team-barrier counter of gram_engine_multi (behind hG, pmaxG, 2 flags)
              double *hG0 = a.scr + 8 + 9 * (long long)a.p + 32 + 4 * NV_GCAP + 8;
              __stcg(reinterpret_cast<unsigned *>(reinterpret_cast<int *>(hG0 + 72) + 2), 0u);
            }
            fast_grid_sync(c);
            const long long tg2 = clock64();
            if (a.multi_ok > 0 && m_act >= a.multi_ok && c.G >= 2) { // a.multi_ok: smallest list length for the team engine (chain CTA + owners)
              if (c.bid < NV_MULTI_TEAM) gram_engine_multi(c, lam, a.maxIter - iter, pass_counter, m_act, Gs, ds);
              if (a.prof && c.bid == 0 && tid == 0) a.prof[10] += clock64() - tg2;
            } else if (c.bid == 0) {
              gram_engine(c, lam, a.maxIter - iter, pass_counter, m_act, Gs, ds);
            }
          } else if (c.bid == 0) {
            active_phase(c, lam, a.maxIter - iter, pass_counter);
          }
          pf[5] += clock64() - t0;
          fast_grid_sync(c);
          const long long np = __ldcg(&bc->npasses);
          if (c.bid != 0) {
            for (int i = tid; i < n; i += NV_T) c.r[i] = __ldcg(a.r + i);
            __syncthreads();
          }
          if (a.kind == CDGPU_LOSS_SQRT) c.rr = shared_sumsq(c);
          iter += np;
          pass_counter += np;
          st.passes += np;
          st.visits += __ldcg(&bc->visits);
          st.accepted += __ldcg(&bc->accepted);
          st.maxH = __ldcg(&bc->maxH);
          conv = __ldcg(&bc->conv) != 0;
          nact_hint = __ldcg(&bc->nact);
          hint_stale = false;
          fast_grid_sync(c); // bc may be rewritten only after everyone has read it
        }
      }
      if (!a.scaled) break;
      // sigma update, lasso.jl:134-140 (every CTA computes the same value from its copy of r)
      st.outer_iters = (int)outer;
      const double snew = sqrt(shared_sumsq(c) / (double)n);
      if (fabs(snew - sigma) / sigma < a.outerTol) break;
      sigma = snew;
    }
    if (a.scaled) st.sigma = sigma;
    // ---- end of this lambda
    if (c.bid == 0 && tid == 0) {
      bc->nact = c.sm->nact;
      __threadfence();
    }
    fast_grid_sync(c);
    const int nnz = __ldcg(&bc->nact);
    nact_hint = nnz;
    hint_stale = false;
    if (!a.accumulate) {
      if (a.colptr && out_off + nnz > a.capacity) status = 1;
      if (c.bid == 0 && status == 0) {
        if (a.colptr) {
          for (int i = tid; i < nnz; i += NV_T) {
            a.rowval[out_off + i] = (long long)a.act[i] + 1;
            a.nzval[out_off + i] = a.actval[i];
          }
          if (tid == 0) a.colptr[li + 1] = out_off + nnz;
        }
        if (tid == 0 && a.stats) a.stats[li] = st;
      }
      out_off += nnz;
      if (status == 0) cols_done = li + 1;
      if (a.max_hat_s >= 0 && nnz > a.max_hat_s) break;
    } else {
      cols_done = li + 1;
    }
  }
  if (c.bid == 0) {
    if (a.accumulate && tid == 0 && a.stats) a.stats[0] = st;
    for (int i = tid; i < n; i += NV_T) a.r[i] = c.r[i];
    if (a.scaled) { // std(f.r) of the returned LassoSolution (lasso.jl:143)
      const double sd = shared_std(c);
      if (tid == 0) a.scr[0] = sd;
    }
    if (tid == 0) {
      if (a.prof) {
        pf[6] = clock64() - t_start;
        for (int i = 0; i < 10; ++i) a.prof[i] = pf[i];
      }
      *a.nact = c.sm->nact;
      a.flag[0] = status;
      a.flag[1] = (int)cols_done;
    }
  }
}

// ------------------------------------------------------------ small kernels --
// initialize!: r = y - X beta over the stored entries (cd_differentiable_function.jl:59-72)
__global__ void naive_init_kernel(const double *X, long long ldx, int n, const double *y, const int *act,
                                  const double *actval, const int *nact, double *r) {
  const int m = *nact;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (int s = 0; s < m; ++s) acc += X[i + (long long)act[s] * ldx] * actval[s];
    r[i] = y[i] - acc;
  }
}
__global__ void dense_iterate_kernel(int p, const int *act, const double *actval, const int *nact, double *beta,
                                     unsigned char *inlist, int phase) {
  const int m = *nact;
  if (phase == 0) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < p; j += gridDim.x * blockDim.x) {
      beta[j] = 0.0;
      inlist[j] = 0;
    }
  } else {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
      beta[act[i]] = actval[i];
      inlist[act[i]] = 1;
    }
  }
}
// a_k = sum_i [w_i] X_ik^2 (one warp per column); sqrt_over_n: _stdX! (utils.jl:127-151)
__global__ void colsq_kernel(const double *X, long long ldx, int n, int p, const double *w, double *out,
                             int sqrt_over_n) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int k = blockIdx.x * wpb + (threadIdx.x >> 5); k < p; k += gridDim.x * wpb) {
    const double *col = X + (long long)k * ldx;
    double s = 0.0;
    if (w)
      for (int i = lane; i < n; i += 32) s = fma(w[i], col[i] * col[i], s);
    else
      for (int i = lane; i < n; i += 32) s = fma(col[i], col[i], s);
    s = warp_sum(s);
    if (lane == 0) out[k] = sqrt_over_n ? sqrt(s / (double)n) : s;
  }
}
// |gradient_k(0)| [/ omega_k] per column, then the max (_findLambdaMax, coordinate_descent.jl:118-149)
__global__ void grad0_kernel(int kind, const double *X, long long ldx, int n, int p, const double *y, const double *w,
                             const double *omega, double *out) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  double ynorm = 1.0;
  if (kind == CDGPU_LOSS_SQRT) { // gradient = -X_j'r / norm(r)   :234-235
    double s = 0.0;
    for (int i = lane; i < n; i += 32) s = fma(y[i], y[i], s);
    ynorm = sqrt(warp_sum(s));
  }
  for (int k = blockIdx.x * wpb + (threadIdx.x >> 5); k < p; k += gridDim.x * wpb) {
    const double *col = X + (long long)k * ldx;
    double s = 0.0;
    if (w)
      for (int i = lane; i < n; i += 32) s = fma(w[i] * col[i], y[i], s);
    else
      for (int i = lane; i < n; i += 32) s = fma(col[i], y[i], s);
    s = warp_sum(s);
    if (lane == 0) {
      double t = kind < 0 ? fabs(s) : fabs(kind == CDGPU_LOSS_SQRT ? s / ynorm : s / (double)n);
      if (omega) t = t / omega[k];
      out[k] = t;
    }
  }
}
__global__ void max_kernel(const double *v, int p, double *out) {
  __shared__ double red[32];
  double m = 0.0;
  for (int j = threadIdx.x; j < p; j += blockDim.x) m = fmax(m, v[j]);
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
    m = warp_max(m);
    if (threadIdx.x == 0) *out = m;
  }
}

} // namespace

int launch_colsq(cdgpu_handle_s *h, const double *X, long long ldx, int n, int p, const double *w, double *out,
                 bool sqrt_over_n) {
  int blocks = min((p + 7) / 8, h->sm_count * 8);
  colsq_kernel<<<max(blocks, 1), 256, 0, h->stream>>>(X, ldx, n, p, w, out, sqrt_over_n ? 1 : 0);
  CUDA_TRY(cudaGetLastError());
  CD_COUNT_LAUNCH(1);
  return CDGPU_OK;
}

// |X_j'y| for every column (the screening statistic of _findLargestCorrelations, utils.jl:96-107)
int launch_abs_xty(cdgpu_handle_s *h, const double *X, long long ldx, int n, int p, const double *y, double *out) {
  int blocks = min((p + 7) / 8, h->sm_count * 8);
  grad0_kernel<<<max(blocks, 1), 256, 0, h->stream>>>(-1, X, ldx, n, p, y, nullptr, nullptr, out);
  CUDA_TRY(cudaGetLastError());
  CD_COUNT_LAUNCH(1);
  return CDGPU_OK;
}

int launch_lambda_max_naive(cdgpu_handle_s *h, int kind, const double *X, long long ldx, int n, int p, const double *y,
                            const double *w, const double *omega, double *scr, double *out) {
  int blocks = min((p + 7) / 8, h->sm_count * 8);
  grad0_kernel<<<max(blocks, 1), 256, 0, h->stream>>>(kind, X, ldx, n, p, y, w, omega, scr);
  max_kernel<<<1, 1024, 0, h->stream>>>(scr, p, out);
  CUDA_TRY(cudaGetLastError());
  CD_COUNT_LAUNCH(2);
  return CDGPU_OK;
}

int launch_naive_init(cdgpu_handle_s *h, const NaiveArgs &a) {
  naive_init_kernel<<<(a.n + 255) / 256, 256, 0, h->stream>>>(a.X, a.ldx, a.n, a.y, a.act, a.actval, a.nact, a.r);
  dense_iterate_kernel<<<(a.p + 255) / 256, 256, 0, h->stream>>>(a.p, a.act, a.actval, a.nact, a.beta, a.inlist, 0);
  dense_iterate_kernel<<<32, 256, 0, h->stream>>>(a.p, a.act, a.actval, a.nact, a.beta, a.inlist, 1);
  CUDA_TRY(cudaGetLastError());
  CD_COUNT_LAUNCH(3);
  return CDGPU_OK;
}

// does the residual (and w) of an n-row problem fit the naive kernel's shared memory?
bool naive_fits(long long n, bool has_w) {
  const size_t rbytes = (size_t)((n + 1) & ~1ll) * sizeof(double) * (has_w ? 2 : 1) + (sizeof(NSmem) + 15) / 16 * 16;
  return rbytes <= (size_t)227 * 1024;
}

int launch_naive_path(cdgpu_handle_s *h, const NaiveArgs &a) {
  static bool attr_done[64] = {false}; // function attributes are per device (context), not per process
  const size_t max_dyn = 227 * 1024;
  // tall sqrt-lasso problems (r does not fit one CTA's shared memory): rows dealt over the grid, tall_sweep.cu
  if (a.kind == CDGPU_LOSS_SQRT && !a.scaled && (!naive_fits(a.n, false) || getenv("CDGPU_FORCE_TALL")))
    return launch_tall_sqrt(h, a);
  const bool known = h->device >= 0 && h->device < 64;
  if (!known || !attr_done[h->device]) {
    CUDA_TRY(cudaFuncSetAttribute(naive_path_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_dyn));
    CUDA_TRY(cudaFuncSetAttribute(naive_path_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_dyn));
    if (known) attr_done[h->device] = true;
  }
  // shared memory: r (and w) + the state of the covariance-form active engine, whose capacity shrinks
  // (down to 0 = disabled, CTA 0 then runs active passes column by column) as n grows
  const size_t rbytes = (size_t)((a.n + 1) & ~1) * sizeof(double) * (a.w ? 2 : 1) + (sizeof(NSmem) + 15) / 16 * 16;
  auto engine_bytes = [](int gcap) {
    return gcap ? chain::STAGE_DOUBLES * sizeof(double) + (size_t)gcap * (2 * sizeof(double) + 2 * sizeof(int) + 2 * sizeof(unsigned short))
                : (size_t)0;
  };
  int gcap = a.gram ? a.gram_cap : 0;
  while (gcap >= 256 && rbytes + engine_bytes(gcap) > max_dyn) gcap >>= 1;
  if (gcap < 256) gcap = 0;
  const size_t dyn = rbytes + engine_bytes(gcap);
  if (dyn > max_dyn)
    return cdgpu_set_error(CDGPU_ECAP,
                           "naive-form sweep keeps r%s in shared memory: n = %d exceeds the %zu-byte limit; use the "
                           "covariance form (cdgpu_gram_create) for tall problems",
                           a.w ? " and w" : "", a.n, max_dyn);
  // grid: one CTA per SM, fewer when there are not enough columns to feed 16 warps each
  int G = min(h->sm_count, (a.p + NV_W - 1) / NV_W);
  if (const char *env = getenv("CDGPU_NAIVE_GRID")) {
    int v = atoi(env);
    if (v >= 1 && v <= h->sm_count) G = v;
  }
  int occ = 0;
  // the instance with dense mode and super-windows for problems whose X stays in L2 (see full_pass)
  const bool dense_kernel = a.dense && (long long)a.n * a.p * 8 <= (64ll << 20);
  const void *kfn = dense_kernel ? (const void *)naive_path_kernel<true> : (const void *)naive_path_kernel<false>;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, NV_T, dyn));
  if (occ < 1) return cdgpu_set_error(CDGPU_ECUDA, "naive sweep kernel does not fit on an SM");
  G = max(1, min(G, occ * h->sm_count));
  // chunk: up to 8 columns per warp per round, at most ~48 MB of columns so a re-evaluation hits L2
  long long cols_l2 = (48ll << 20) / ((long long)a.n * 8);
  long long CH = min((long long)G * NV_W * 8, max((long long)G * NV_W * 4, cols_l2));
  CH = max(1ll, min(CH, (long long)a.p));
  if (const char *env = getenv("CDGPU_NAIVE_CHUNK")) {
    long long v = atoll(env);
    if (v >= 1) CH = min(v, (long long)a.p);
  }
  // scratch doubles: [0,8) misc | (8p free) | bc 16 | compaction staging p | ... | tail: the 4 result buffers of the
  // full-pass rounds, 4 * CH entries of 32 B <= 16p doubles (handle_common_alloc)
  HEntry *hbuf = reinterpret_cast<HEntry *>(a.scr + cd_scr_tail((size_t)a.p, (size_t)a.n));
  NBcast *bc = reinterpret_cast<NBcast *>(a.scr + 8 + 8 * (long long)a.p);
  int ch = (int)CH;
  void *args[] = {(void *)&a, (void *)&ch, (void *)&hbuf, (void *)&bc, (void *)&gcap};
  CUDA_TRY(cudaLaunchCooperativeKernel(kfn, dim3(G), dim3(NV_T), args, dyn, h->stream));
  CD_COUNT_LAUNCH(1);
  return CDGPU_OK;
}
