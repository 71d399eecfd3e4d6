// vc_batch.cu — K4: batched varying-coefficient lasso, locpolyl1 with refit=false
// (src/varying_coefficient_lasso.jl:30-79): for every grid point z0 one kernel-weighted
// local-polynomial lasso  min sum_i w_i (y_i - eX_i'b)^2/(2n) + lambda0 sum_k sd_k |b_k|,
//   w_i = K_h(z_i, z0) (:17-21),  eX[i,(j,l)] = X[i,j] (z_i - z0)^l (:550-569),
//   sd_k = sqrt(sum_i w_i eX_ik^2 / n) (utils.jl:140-151),
// solved by the reference's CD loop on CDWeightedLSLoss (cd_differentiable_function.jl:165-194).
//
// B200 design.  DEFAULT (ep <= 256): the MOMENT FORM further down — all local Gram matrices of all grid points are
// one FP64 tensor-core GEMM (Z'V, gram_dmma.cu) and each local lasso is a covariance-form CD solved by ONE WARP
// (vc_cov_kernel), with the refit (cdgpu_vc_solve_refit) and the leave-one-out scaled-lasso problems of
// lvocv_locpolyl1 (cdgpu_vc_lvocv) in the same kernel.  RESIDUAL FORM (CDGPU_VC_FORM=naive, or ep > 256), first
// in this file: ONE WARP (n <= 512) or ONE CTA PER GRID POINT, everything per-problem (w, z - z0, r, column
// norms, iterate, active list) in registers / shared memory; the expanded n x p(d+1) design is never
// materialised — a column is X[:,j] (shared by all problems, L1/L2 resident: 8np bytes total) times a power of
// (z - z0).  No inter-CTA communication at all; grid points are sharded over GPUs by [m_begin, m_end).
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include <type_traits>

#include "common.cuh"

#define API extern "C" __attribute__((visibility("default")))

namespace {

// threads per local problem: a single warp when n is small (no block-wide barriers on the per-step
// chain, ~14 problems resident per SM), more warps for longer columns

struct VcArgs {
  const double *X;
  long long ldx;
  int n, p, degree;
  const double *z, *y, *zgrid;
  int g0, g1; // grid points [g0, g1)
  int kernel_kind;
  double bandwidth, lambda0;
  long long maxIter;
  double optTol;
  int randomize;
  unsigned long long seed;
  double *out; // ep x m, column g at out + g*ep
  DevStats *stats;
};

constexpr int VC_WMAX = 8;
struct VSm {
  double red[VC_WMAX];
  double cand_h[VC_WMAX], cand_nw[VC_WMAX];
  int cand_app[VC_WMAX];
  int nact, flag, nonapp;
  int s2[2];
};

__device__ __forceinline__ double ipow(double x, int l) {
  double v = 1.0;
  for (int i = 0; i < l; ++i) v *= x;
  return v;
}

template <int VC_W>
__device__ __forceinline__ double vblock_sum(VSm *sm, double v) {
  v = warp_sum(v);
  if (VC_W == 1) return v;
  if ((threadIdx.x & 31) == 0) sm->red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int i = 0; i < VC_W; ++i) t += sm->red[i];
  __syncthreads();
  return t;
}

template <int VC_T>
__global__ void __launch_bounds__(VC_T) vc_kernel(const VcArgs a) {
  constexpr int VC_W = VC_T / 32;
  extern __shared__ __align__(16) unsigned char raw[];
  const int n = a.n, dg = a.degree + 1, ep = a.p * dg;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  VSm *sm = reinterpret_cast<VSm *>(raw);
  double *sw = reinterpret_cast<double *>(raw + (sizeof(VSm) + 15) / 16 * 16);
  double *sdz = sw + n, *sr = sdz + n;
  double *sa = sr + n;       // a_k = sum w eX_k^2
  double *sbeta = sa + ep;   // dense iterate
  double *sval = sbeta + ep; // values in list order
  double *stmpd = sval + ep; // compaction staging
  int *sact = reinterpret_cast<int *>(stmpd + ep);
  int *snewpos = sact + ep, *snonapp = snewpos + ep, *stmpi = snonapp + ep; // stmpi: 5*ep
  unsigned char *sin = reinterpret_cast<unsigned char *>(stmpi + 5 * ep);

  for (int g = a.g0 + blockIdx.x; g < a.g1; g += gridDim.x) {
    const double z0 = a.zgrid[g];
    __syncthreads();
    for (int i = tid; i < n; i += VC_T) {
      const double zi = a.z[i];
      double w;
      if (a.kernel_kind == CDGPU_KERNEL_GAUSSIAN) {
        const double d = zi - z0;
        w = exp(-(d * d) / a.bandwidth) / a.bandwidth;
      } else {
        const double u = (zi - z0) / a.bandwidth;
        w = fabs(u) >= 1.0 ? 0.0 : 0.75 * (1.0 - u * u) / a.bandwidth;
      }
      sw[i] = w;
      sdz[i] = zi - z0;
      sr[i] = a.y[i]; // r = y - eX*0
    }
    for (int k = tid; k < ep; k += VC_T) {
      sbeta[k] = 0.0;
      sin[k] = 0;
    }
    if (tid == 0) sm->nact = 0;
    __syncthreads();
    // column norms, one warp per expanded column
    for (int k = warp; k < ep; k += VC_W) {
      const int j = k / dg, l = k - j * dg;
      const double *col = a.X + (long long)j * a.ldx;
      double s = 0.0;
      for (int i = lane; i < n; i += 32) {
        const double e = __ldg(col + i) * ipow(sdz[i], l);
        s = fma(sw[i], e * e, s);
      }
      s = warp_sum(s);
      if (lane == 0) sa[k] = s;
    }
    __syncthreads();

    DevStats st;
    st.passes = st.full_passes = st.visits = st.accepted = 0;
    st.maxH = 0.0;
    st.converged = 0;
    st.outer_iters = 0;
    st.sigma = 0.0;
    const bool ordered = a.randomize == 0;
    unsigned long long pass_counter = 0;
    bool conv = true;
    long long iter = 0;
    while (iter < a.maxIter) {
      double maxH = 0.0;
      iter += 1;
      st.passes += 1;
      if (conv) { // ---- full pass, speculative in rounds of VC_W coordinates
        st.full_passes += 1;
        st.visits += ep;
        const PermKey pk = cd_perm_key((uint32_t)ep, a.seed, pass_counter);
        const int m_old = sm->nact;
        if (tid == 0) sm->nonapp = 0;
        int pos = 0;
        while (pos < ep) {
          const int myq = pos + warp;
          double h = 0.0, nw = 0.0;
          int k = -1, app = 1;
          if (myq < ep) {
            k = ordered ? myq : (int)cd_perm(pk, (uint32_t)myq);
            const int j = k / dg, l = k - j * dg;
            const double *col = a.X + (long long)j * a.ldx;
            double d = 0.0;
            for (int i = lane; i < n; i += 32) d = fma(sr[i] * (__ldg(col + i) * ipow(sdz[i], l)), sw[i], d);
            d = warp_sum(d);
            const double ak = sa[k], old = sbeta[k];
            const double v = __dadd_rn(old, d / ak);
            const double thr = __dmul_rn(__dmul_rn((double)n / ak, a.lambda0), sqrt(ak / (double)n));
            nw = cd_shrink(v, thr);
            h = nw - old;
            app = (v != 0.0 || sin[k]) ? 1 : 0; // `x[k] += b/a` appends a non-member iff the sum is non-zero
          }
          if (lane == 0) {
            sm->cand_h[warp] = h;
            sm->cand_nw[warp] = nw;
            sm->cand_app[warp] = app;
          }
          __syncthreads();
          int first = -1;
#pragma unroll
          for (int q = VC_W - 1; q >= 0; --q)
            if (sm->cand_h[q] != 0.0) first = q;
          if (tid == 0) { // finalised visits of this round that the reference would not have appended (rare)
            const int qend = first < 0 ? VC_W : first;
            for (int q = 0; q < qend; ++q)
              if (!sm->cand_app[q] && pos + q < ep) snonapp[sm->nonapp++] = ordered ? pos + q : (int)cd_perm(pk, (uint32_t)(pos + q));
          }
          if (first < 0) {
            pos += VC_W;
            __syncthreads();
            continue;
          }
          const double hh = sm->cand_h[first], nn = sm->cand_nw[first];
          const int kq = pos + first;
          const int kk = ordered ? kq : (int)cd_perm(pk, (uint32_t)kq);
          __syncthreads();
          if (tid == 0) {
            sbeta[kk] = nn;
            if (!sin[kk]) {
              sin[kk] = 1;
              sact[sm->nact] = kk;
              sm->nact += 1;
            }
          }
          {
            const int j = kk / dg, l = kk - j * dg;
            const double *col = a.X + (long long)j * a.ldx;
            for (int i = tid; i < n; i += VC_T) sr[i] = __dsub_rn(sr[i], __dmul_rn(__ldg(col + i) * ipow(sdz[i], l), hh));
          }
          maxH = fmax(maxH, fabs(hh));
          st.accepted += 1;
          pos = kq + 1;
          __syncthreads();
        }
        // list update: values, then the reference's post-dropzeros! order (common.cuh: cd_compact_list)
        const int mnow = sm->nact, nna = sm->nonapp;
        for (int i = tid; i < mnow; i += VC_T) sval[i] = sbeta[sact[i]];
        for (int e = m_old + tid; e < mnow; e += VC_T) {
          const int k = sact[e];
          const int vis = ordered ? k : (int)cd_perm_inv(pk, (uint32_t)k);
          int before = 0;
          for (int j = 0; j < m_old; ++j) before += (ordered ? sact[j] : (int)cd_perm_inv(pk, (uint32_t)sact[j])) < vis;
          for (int j = 0; j < nna; ++j) before += (ordered ? snonapp[j] : (int)cd_perm_inv(pk, (uint32_t)snonapp[j])) < vis;
          snewpos[e - m_old] = m_old + vis - before;
        }
        __syncthreads();
        cd_compact_list<VC_T>(sact, sval, m_old, mnow, snewpos, sin, stmpi, stmpd, sm->s2);
        if (tid == 0) sm->nact = sm->s2[0];
        __syncthreads();
      } else { // ---- active-set pass: sequential chain, block-wide dot per step
        const int m = sm->nact;
        st.visits += m;
        const PermKey pkm = cd_perm_key((uint32_t)max(m, 1), a.seed, pass_counter);
        for (int s = 0; s < m; ++s) {
          const int i_ = ordered ? s : (int)cd_perm(pkm, (uint32_t)s);
          const int k = sact[i_];
          const int j = k / dg, l = k - j * dg;
          const double *col = a.X + (long long)j * a.ldx;
          double d = 0.0;
          for (int i = tid; i < n; i += VC_T) d = fma(sr[i] * (__ldg(col + i) * ipow(sdz[i], l)), sw[i], d);
          d = vblock_sum<VC_W>(sm, d);
          const double ak = sa[k], old = sval[i_];
          const double v = __dadd_rn(old, d / ak);
          const double thr = __dmul_rn(__dmul_rn((double)n / ak, a.lambda0), sqrt(ak / (double)n));
          const double nw = cd_shrink(v, thr);
          const double h = nw - old;
          __syncthreads();
          if (tid == 0) {
            sval[i_] = nw;
            sbeta[k] = nw;
          }
          if (h != 0.0) {
            for (int i = tid; i < n; i += VC_T) sr[i] = __dsub_rn(sr[i], __dmul_rn(__ldg(col + i) * ipow(sdz[i], l), h));
            st.accepted += 1;
          }
          maxH = fmax(maxH, fabs(h));
          __syncthreads();
        }
      }
      pass_counter += 1;
      // ---- dropzeros! (after a full pass the list was already compacted above)
      if (tid == 0) {
        int nn = sm->nact, i = 0;
        while (i < nn) {
          if (sval[i] == 0.0) {
            sin[sact[i]] = 0;
            if (i != nn - 1) {
              sval[i] = sval[nn - 1];
              sact[i] = sact[nn - 1];
            }
            nn -= 1;
          } else {
            i += 1;
          }
        }
        sm->nact = nn;
      }
      __syncthreads();
      st.maxH = maxH;
      const bool prev = conv;
      conv = maxH < a.optTol;
      if (prev && conv) {
        st.converged = 1;
        break;
      }
    }
    double *col = a.out + (long long)g * ep;
    for (int k = tid; k < ep; k += VC_T) col[k] = sbeta[k];
    if (tid == 0 && a.stats) a.stats[g] = st;
  }
}


// ------------------------------------------------------------------------------------------------
// One WARP per local problem, for n <= 32*NR: w, z - z0 and the residual live in REGISTERS (NR
// values per lane), a visit is NR independent loads of the shared X column (L1/L2 resident), a
// shuffle reduction and, when the coordinate moves, NR register FMAs.  No block barrier on the
// per-step chain, ~10 KB of shared memory per problem (only the ep-sized iterate/list state), so
// ~12 problems are resident per SM and X keeps most of L1.
template <int NR>
__global__ void __launch_bounds__(32) vc_warp_kernel(const VcArgs a) {
  extern __shared__ __align__(16) unsigned char raw[];
  const int n = a.n, dg = a.degree + 1, ep = a.p * dg, lane = threadIdx.x;
  double *sa = reinterpret_cast<double *>(raw);
  double *sbeta = sa + ep, *sval = sbeta + ep, *stmpd = sval + ep;
  int *sact = reinterpret_cast<int *>(stmpd + ep);
  int *snewpos = sact + ep, *snonapp = snewpos + ep, *stmpi = snonapp + ep, *s2 = stmpi + 5 * ep;
  unsigned char *sin = reinterpret_cast<unsigned char *>(s2 + 4);
  const bool ordered = a.randomize == 0;

  for (int g = a.g0 + blockIdx.x; g < a.g1; g += gridDim.x) {
    const double z0 = a.zgrid[g];
    double w[NR], dz[NR], r[NR];
#pragma unroll
    for (int t = 0; t < NR; ++t) {
      const int i = lane + 32 * t;
      w[t] = dz[t] = r[t] = 0.0;
      if (i < n) {
        const double zi = a.z[i];
        if (a.kernel_kind == CDGPU_KERNEL_GAUSSIAN) {
          const double d = zi - z0;
          w[t] = exp(-(d * d) / a.bandwidth) / a.bandwidth;
        } else {
          const double u = (zi - z0) / a.bandwidth;
          w[t] = fabs(u) >= 1.0 ? 0.0 : 0.75 * (1.0 - u * u) / a.bandwidth;
        }
        dz[t] = zi - z0;
        r[t] = a.y[i];
      }
    }
    __syncwarp();
    for (int k = lane; k < ep; k += 32) {
      sbeta[k] = 0.0;
      sin[k] = 0;
    }
    // expanded column k = (j, l): e_t = X[i, j] * dz^l; returns sum_t w e r (dot) or sum_t w e^2 (norm)
    auto column = [&](int k, double (&e)[NR]) {
      const int j = k / dg, l = k - j * dg;
      const double *col = a.X + (long long)j * a.ldx;
#pragma unroll
      for (int t = 0; t < NR; ++t) {
        const int i = lane + 32 * t;
        e[t] = i < n ? __ldg(col + i) : 0.0;
      }
      for (int q = 0; q < l; ++q) {
#pragma unroll
        for (int t = 0; t < NR; ++t) e[t] *= dz[t];
      }
    };
    for (int k = 0; k < ep; ++k) {
      double e[NR];
      column(k, e);
      double s = 0.0;
#pragma unroll
      for (int t = 0; t < NR; ++t) s = fma(w[t], e[t] * e[t], s);
      s = warp_sum(s);
      if (lane == 0) sa[k] = s;
    }
    int nact = 0;
    __syncwarp();

    DevStats st;
    st.passes = st.full_passes = st.visits = st.accepted = 0;
    st.maxH = 0.0;
    st.converged = 0;
    st.outer_iters = 0;
    st.sigma = 0.0;
    unsigned long long pass_counter = 0;
    bool conv = true;
    long long iter = 0;
    // one visit of coordinate k whose current value is `old`; returns h (uniform across the warp)
    auto visit = [&](int k, double old, double &nw, bool &tnz) -> double {
      double e[NR];
      column(k, e);
      double d = 0.0;
#pragma unroll
      for (int t = 0; t < NR; ++t) d = fma(r[t] * e[t], w[t], d);
      d = warp_sum(d);
      const double ak = sa[k];
      const double v = __dadd_rn(old, d / ak);
      const double thr = __dmul_rn(__dmul_rn((double)n / ak, a.lambda0), sqrt(ak / (double)n));
      nw = cd_shrink(v, thr);
      tnz = v != 0.0;
      const double h = nw - old;
      if (h != 0.0) {
#pragma unroll
        for (int t = 0; t < NR; ++t) r[t] = __dsub_rn(r[t], __dmul_rn(e[t], h));
      }
      return h;
    };
    while (iter < a.maxIter) {
      double maxH = 0.0;
      iter += 1;
      st.passes += 1;
      if (conv) { // ---- full pass
        st.full_passes += 1;
        st.visits += ep;
        const PermKey pk = cd_perm_key((uint32_t)ep, a.seed, pass_counter);
        const int m_old = nact;
        int nna = 0;
        for (int q = 0; q < ep; ++q) {
          const int k = ordered ? q : (int)cd_perm(pk, (uint32_t)q);
          const double old = sbeta[k];
          double nw;
          bool tnz;
          const double h = visit(k, old, nw, tnz);
          const bool member = sin[k] != 0;
          __syncwarp();
          if (!tnz && !member) { // not appended by `x[k] += b/a` (rare)
            if (lane == 0) snonapp[nna] = k;
            nna += 1;
          }
          if (h != 0.0) {
            if (lane == 0) {
              sbeta[k] = nw;
              if (!member) {
                sin[k] = 1;
                sact[nact] = k;
              }
            }
            if (!member) nact += 1;
            maxH = fmax(maxH, fabs(h));
            st.accepted += 1;
          }
          __syncwarp();
        }
        for (int i = lane; i < nact; i += 32) sval[i] = sbeta[sact[i]];
        for (int e = m_old + lane; e < nact; e += 32) {
          const int k = sact[e];
          const int vis = ordered ? k : (int)cd_perm_inv(pk, (uint32_t)k);
          int before = 0;
          for (int j = 0; j < m_old; ++j) before += (ordered ? sact[j] : (int)cd_perm_inv(pk, (uint32_t)sact[j])) < vis;
          for (int j = 0; j < nna; ++j) before += (ordered ? snonapp[j] : (int)cd_perm_inv(pk, (uint32_t)snonapp[j])) < vis;
          snewpos[e - m_old] = m_old + vis - before;
        }
        __syncwarp();
        cd_compact_list<32>(sact, sval, m_old, nact, snewpos, sin, stmpi, stmpd, s2);
        nact = s2[0];
        __syncwarp();
      } else { // ---- active-set pass
        const int m = nact;
        st.visits += m;
        const PermKey pkm = cd_perm_key((uint32_t)max(m, 1), a.seed, pass_counter);
        for (int s = 0; s < m; ++s) {
          const int i_ = ordered ? s : (int)cd_perm(pkm, (uint32_t)s);
          const int k = sact[i_];
          const double old = sval[i_];
          double nw;
          bool tnz;
          const double h = visit(k, old, nw, tnz);
          __syncwarp();
          if (lane == 0) {
            sval[i_] = nw;
            sbeta[k] = nw;
          }
          if (h != 0.0) st.accepted += 1;
          maxH = fmax(maxH, fabs(h));
          __syncwarp();
        }
        cd_compact_list<32>(sact, sval, m, m, snewpos, sin, stmpi, stmpd, s2); // dropzeros!
        nact = s2[0];
        __syncwarp();
      }
      pass_counter += 1;
      st.maxH = maxH;
      const bool prev = conv;
      conv = maxH < a.optTol;
      if (prev && conv) {
        st.converged = 1;
        break;
      }
    }
    double *col = a.out + (long long)g * ep;
    for (int k = lane; k < ep; k += 32) col[k] = sbeta[k];
    if (lane == 0 && a.stats) a.stats[g] = st;
    __syncwarp();
  }
}


// ------------------------------------------------------------------------------------------------
// MOMENT FORM (default for ep <= 256): the covariance form of every local problem at once.
// With eX[i,(j,l)] = X[i,j] dz_i^l the weighted Gram of grid point g is
//     A_g[(j,l),(j',l')] = sum_i w_gi dz_gi^(l+l') X[i,j] X[i,j'] / n = M_{g,l+l'}[j,j'],
// i.e. 2d+1 weighted p x p moment matrices per grid point, and ALL of them for ALL grid points are
// ONE dense FP64 contraction  C = Z'V / n  with  Z[i,(j>=j')] = X[i,j] X[i,j'] | Z[i,P2+j] = X[i,j] y_i
// (n x (p(p+1)/2 + p)) and V[i,(g,q)] = w_gi dz_gi^q (n x m(2d+1)) — run on the FP64 tensor cores by
// the DMMA kernel of gram_dmma.cu (launch_gemm_tn).  The local lasso is then CDQuadraticLoss
// (cd_differentiable_function.jl:324-348) on A_g, b_g = -eX'W y/n with omega_k = sqrt(A_kk)
// (utils.jl:140-151) — the same minimiser as the reference's CDWeightedLSLoss form (:165-194);
// iterates differ at rounding level.  ONE WARP PER GRID POINT: lane owns coordinates lane, lane+32, ...
// with (A x)_t, beta_t and the per-coordinate constants in registers; a full pass is the speculative
// first-mover scan (non-moving coordinates cost no memory traffic), an accepted step reads one
// column of A_g through the moment blocks (L1/L2 resident).
struct VcCovArgs {
  const double *C; // moment blocks: problem g at C + (g - g0) * NQ * ldc, block q at + q * ldc
  long long ldc;
  int p, degree, ep, P2;
  int g0, g1;
  double lambda0;
  long long maxIter;
  double optTol;
  int randomize;
  unsigned long long seed;
  double *out, *outR; // outR: refitted coefficients (null: no refit)
  DevStats *stats;
  int *counter; // dynamic work distribution
  // lvocv_locpolyl1: problem g = (bandwidth g / n, left-out observation g % n)
  int lvo, n;
  const double *X, *y; // device copies (n x p, ldx) for the prediction of the left-out response
  long long ldx;
  double *lvo_err;     // squared prediction error per problem
  unsigned long long *prof; // optional [8]: summed warp cycles: full passes, active chain, list compaction, phase open, phase close, total
  double *gscr; // per-warp scratch for the compact active Gram: (grid * VCW) x MC x MC doubles, MC = ep rounded up to even
};

constexpr int VCW = 1;      // warps (local problems) per CTA (one: shared memory then packs 11 problems per SM)
constexpr int VC_RING = 4;  // columns of the compact active Gram in flight to shared memory per warp
__host__ __device__ inline size_t vc_cov_warp_bytes(int ep, int nu) { // nu: the kernel instance's slots per lane
  const int RS = 32 * nu; // ring stage stride: a whole number of 32-lane rows, so no lane ever clamps
  return ((size_t)(VC_RING * RS + 8 * ep) * sizeof(double) + (size_t)(6 * ep + 4) * sizeof(int) + (size_t)ep + 15) / 16 * 16;
}

__global__ void vc_build_z_kernel(const double *__restrict__ X, long long ldx, int n, int p, const double *__restrict__ y,
                                  double *Z, long long ldz) {
  const int P2 = p * (p + 1) / 2;
  const int col = blockIdx.x;
  if (col >= P2 + p) { // y.^2 and the constant 1: sum w y^2 and sum w for the sigma of the scaled-lasso loops
    for (int i = threadIdx.x; i < n; i += blockDim.x) Z[i + (long long)col * ldz] = col == P2 + p ? y[i] * y[i] : 1.0;
    return;
  }
  int j, jp;
  if (col < P2) {
    j = (int)((sqrt(8.0 * (double)col + 1.0) - 1.0) * 0.5);
    while (j * (j + 1) / 2 > col) --j;
    while ((j + 1) * (j + 2) / 2 <= col) ++j;
    jp = col - j * (j + 1) / 2;
  } else {
    j = col - P2;
    jp = -1;
  }
  const double *cj = X + (long long)j * ldx, *cjp = jp >= 0 ? X + (long long)jp * ldx : y;
  for (int i = threadIdx.x; i < n; i += blockDim.x) Z[i + (long long)col * ldz] = cj[i] * cjp[i];
}

// lvo (harr != null): problem g = (bandwidth harr[g / n], z0 = z[g % n]) with the weight of observation g % n zeroed
__global__ void vc_build_v_kernel(const double *__restrict__ z, const double *__restrict__ zgrid, int n, int g0, int g1,
                                  int nq, int kernel_kind, double bandwidth, double *V, long long ldv,
                                  const double *__restrict__ harr) {
  const int g = g0 + blockIdx.x;
  if (g >= g1) return;
  const double z0 = harr ? z[g % n] : zgrid[g];
  if (harr) bandwidth = harr[g / n];
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double zi = z[i];
    double w;
    if (kernel_kind == CDGPU_KERNEL_GAUSSIAN) { // varying_coefficient_lasso.jl:17
      const double d = zi - z0;
      w = exp(-(d * d) / bandwidth) / bandwidth;
    } else { // :18-21
      const double u = (zi - z0) / bandwidth;
      w = fabs(u) >= 1.0 ? 0.0 : 0.75 * (1.0 - u * u) / bandwidth;
    }
    if (harr && i == g % n) w = 0.0; // w[i] = zero(T), varying_coefficient_lasso.jl:108
    const double dz = zi - z0;
    double v = w;
    for (int q = 0; q < nq; ++q) {
      V[i + (long long)((g - g0) * nq + q) * ldv] = v;
      v *= dz;
    }
  }
}

template <int OFF>
__device__ __forceinline__ void vc_cp16(unsigned dst, const char *src) {
  asm volatile("cp.async.cg.shared.global [%0+%2], [%1+%2], 16;" ::"r"(dst), "l"(src), "n"(OFF) : "memory");
}

template <int NU, bool LVO>
__global__ void __launch_bounds__(VCW * 32, 11) vc_cov_kernel(const VcCovArgs a) {
  extern __shared__ __align__(16) unsigned char raw[];
  const int ep = a.ep, dg = a.degree + 1, nq = 2 * a.degree + 1, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int MC = (ep + 1) & ~1, RS = 32 * NU; // scratch leading dimension bound, ring stage stride
  unsigned char *base = raw + warp * vc_cov_warp_bytes(ep, NU);
  double *ring = reinterpret_cast<double *>(base);       // VC_RING prefetched columns of the compact active Gram
  double *sbeta = ring + VC_RING * RS;                   // dense beta (after a full pass / at phase start)
  double *sval = sbeta + ep, *stmpd = sval + ep;         // list-order values, scratch
  double *sAx = stmpd + ep, *scc = sAx + ep, *sai = scc + ep, *sth = sai + ep; // dense (A x) and constants
  double *sAxE = sth + ep;                               // (A x) of the active entries of a phase, by snapshot entry
  int *sact = reinterpret_cast<int *>(sAxE + ep);
  int *snewpos = sact + ep, *sact0 = snewpos + ep, *stmpi = sact0 + ep, *s2 = stmpi + 2 * ep, *spos = s2 + 4;
  int *scmp = reinterpret_cast<int *>(ring); // cd_compact_list's 5*m ints: the column ring is idle whenever a list is compacted
  unsigned char *sin = reinterpret_cast<unsigned char *>(spos + ep);
  double *Gw = a.gscr + ((long long)blockIdx.x * VCW + warp) * (long long)MC * MC; // this warp's compact Gram scratch
  const bool ordered = a.randomize == 0;
  constexpr unsigned NONE = 0xffffffffu;

  int tj[NU], tl[NU], ttri[NU];
#pragma unroll
  for (int u = 0; u < NU; ++u) {
    const int t = lane + 32 * u;
    tj[u] = t / dg;
    tl[u] = t - tj[u] * dg;
    ttri[u] = tj[u] * (tj[u] + 1) / 2;
  }

  for (;;) {
    int g = 0;
    if (lane == 0) g = a.g0 + atomicAdd(a.counter, 1);
    g = __shfl_sync(0xffffffffu, g, 0);
    if (g >= a.g1) break;
    const double *Cg = a.C + (long long)(g - a.g0) * nq * a.ldc;
    // element (k1, k2) of A_g through the moment blocks
    auto moment = [&](int k1, int k2) -> double {
      const int j1 = k1 / dg, l1 = k1 - j1 * dg, j2 = k2 / dg, l2 = k2 - j2 * dg;
      const int pk_ = j1 >= j2 ? j1 * (j1 + 1) / 2 + j2 : j2 * (j2 + 1) / 2 + j1;
      return __ldg(Cg + (long long)(l1 + l2) * a.ldc + pk_);
    };
    // state: COORDINATE layout outside a phase (slot u <-> coordinate lane + 32u), ENTRY layout inside one
    // (slot u <-> snapshot entry lane + 32u of the phase's active list)
    double Ax[NU], be[NU], cc[NU], ai[NU], th[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      const int t = lane + 32 * u;
      Ax[u] = be[u] = cc[u] = th[u] = 0.0;
      ai[u] = 0.0;
      if (t < ep) {
        const double att = __ldg(Cg + (long long)(2 * tl[u]) * a.ldc + ttri[u] + tj[u]);
        ai[u] = 1.0 / att;
        th[u] = __dmul_rn(__dmul_rn(ai[u], a.lambda0), sqrt(att)); // lambda0 * omega_k / A_kk, omega_k = sqrt(A_kk)
        cc[u] = -__ldg(Cg + (long long)tl[u] * a.ldc + a.P2 + tj[u]);
        scc[t] = cc[u];
        sai[t] = ai[u];
        sth[t] = th[u];
      }
    }
    for (int k = lane; k < ep; k += 32) sin[k] = 0;
    int nact = 0;
    __syncwarp();

    // coordinate layout: Ax[u] += A_g[t_u, k] * h for the mover k = (kj, kl)
    auto apply = [&](int k, double h) {
      const int kj = k / dg, kl = k - kj * dg, ktri = kj * (kj + 1) / 2;
      double gv[NU];
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        const int t = lane + 32 * u;
        const int pk_ = kj >= tj[u] ? ktri + tj[u] : ttri[u] + kj;
        gv[u] = t < ep ? __ldg(Cg + (long long)(kl + tl[u]) * a.ldc + pk_) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < NU; ++u) Ax[u] = __dadd_rn(Ax[u], __dmul_rn(gv[u], h));
    };

    // ---- phases of consecutive active-set passes.  Active-set passes only need (A x) on the active set, so a
    // phase (a) gathers the compact m0 x m0 block A_g[act0, act0] once into this warp's scratch (L2 resident,
    // column-contiguous), (b) moves the state of the active entries into the entry layout, (c) runs the chain
    // with the column of the next VC_RING-1 steps already in flight to shared memory (cp.async), so a step is:
    // owner evaluates, one shuffle broadcasts h, every lane updates its entries — no memory latency on the
    // dependent path, and (d) at the end brings (A x) of the other coordinates up to date from the change of beta.
    bool in_phase = false;
    int m0 = 0, ldw = 0;
    auto phase_open = [&](int m) {
      m0 = m;
      ldw = (m + 1) & ~1;
#pragma unroll
      for (int u = 0; u < NU; ++u)
        if (lane + 32 * u < ep) {
          sbeta[lane + 32 * u] = be[u]; // beta at the start of the phase
          sAx[lane + 32 * u] = Ax[u];   // (A x) at the start of the phase, by coordinate
        }
      for (int k = lane; k < ep; k += 32) spos[k] = -1;
      __syncwarp();
      for (int i = lane; i < m; i += 32) {
        const int k = sact[i];
        sact0[i] = k;
        spos[k] = i;
        sAxE[i] = sAx[k]; // (A x) of the active entries, by snapshot entry, carried from pass to pass
      }
      __syncwarp();
      // compact Gram: 4 columns per iteration so that 4 * ceil(m/32) independent gathers are in flight per lane
      for (int j = 0; j < m; j += 4) {
        int kj[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) kj[c] = sact0[min(j + c, m - 1)];
        for (int i = lane; i < m; i += 32) {
          const int ki = sact0[i];
          double v[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) v[c] = moment(ki, kj[c]);
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (j + c < m) Gw[i + (long long)(j + c) * ldw] = v[c];
        }
      }
      __syncwarp();
      in_phase = true;
    };
    // back to the coordinate layout.  Coordinates in the list take their tracked (A x); every other coordinate
    // (never active in this phase, or dropped during it) gets (A x) at phase start + sum_e A[t, act0_e] (beta_e -
    // beta_e at phase start).
    auto phase_close = [&](int nact_now) {
      for (int k = lane; k < ep; k += 32) stmpd[k] = 0.0; // dense beta now (dropped entries are exactly zero)
      __syncwarp();
      for (int i = lane; i < nact_now; i += 32) {
        const int k = sact[i];
        stmpd[k] = sval[i];
        sAx[k] = sAxE[spos[k]];
      }
      __syncwarp();
      bool inactive[NU];
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        const int t = lane + 32 * u;
        Ax[u] = be[u] = cc[u] = ai[u] = th[u] = 0.0;
        inactive[u] = false;
        if (t < ep) {
          inactive[u] = sin[t] == 0;
          Ax[u] = sAx[t];
          be[u] = stmpd[t];
          cc[u] = scc[t];
          ai[u] = sai[t];
          th[u] = sth[t];
        }
      }
      for (int e0 = 0; e0 < m0; e0 += 2) { // two entries per iteration: twice the loads in flight
        double gv[2][NU], dl[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int k = sact0[min(e0 + c, m0 - 1)];
          dl[c] = e0 + c < m0 ? stmpd[k] - sbeta[k] : 0.0;
          const int kj = k / dg, kl = k - kj * dg, ktri = kj * (kj + 1) / 2;
#pragma unroll
          for (int u = 0; u < NU; ++u) {
            const int pk_ = kj >= tj[u] ? ktri + tj[u] : ttri[u] + kj;
            gv[c][u] = inactive[u] && dl[c] != 0.0 ? __ldg(Cg + (long long)(kl + tl[u]) * a.ldc + pk_) : 0.0;
          }
        }
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int u = 0; u < NU; ++u)
            if (inactive[u]) Ax[u] = fma(gv[c][u], dl[c], Ax[u]);
      }
      __syncwarp();
      in_phase = false;
    };
    // One active-set pass inside a phase; returns max|h|.  The m listed entries are laid out IN VISIT ORDER for the
    // pass: visit position s <-> (lane s & 31, slot s >> 5), so the stepping slot is a compile-time index and the
    // owner is the loop counter: no selects, no uniform loads on the chain.  A step: every lane evaluates its own
    // slot-u entry, lane i's h is broadcast by one shuffle, every lane updates all its entries with the column of
    // the compact Gram that the cp.async ring brought to shared memory VC_RING-1 steps earlier (rows are read
    // through the lane's position -> snapshot-entry map).
    // MS = ceil(m / 32) slots per lane are live in the pass: the per-step work (gathers, multiply-adds) is instantiated
    // for exactly that many, instead of the NU the widest possible list would need
    auto phase_pass = [&](int m, const PermKey &pkm, int &accepted, auto ms_tag) -> double {
      constexpr int MS = decltype(ms_tag)::value;
      unsigned *soff = reinterpret_cast<unsigned *>(stmpi); // byte offset of the column of visit position s in Gw
      int *slist = stmpi + ep;                              // list position i_ of visit position s
      for (int s_ = lane; s_ < m; s_ += 32) {
        const int i_ = ordered ? s_ : (int)cd_perm(pkm, (uint32_t)s_);
        slist[s_] = i_;
        soff[s_] = (unsigned)(spos[sact[i_]] * ldw) * 8u;
      }
      __syncwarp();
      int row[MS]; // snapshot entry (= row of the compact Gram) of this lane's visit positions
#pragma unroll
      for (int u = 0; u < MS; ++u) {
        const int s_ = lane + 32 * u;
        Ax[u] = be[u] = cc[u] = ai[u] = th[u] = 0.0;
        row[u] = 0;
        if (s_ < m) {
          const int i_ = slist[s_], k = sact[i_];
          row[u] = spos[k];
          Ax[u] = sAxE[row[u]];
          be[u] = sval[i_];
          cc[u] = scc[k];
          ai[u] = sai[k];
          th[u] = sth[k];
        }
      }
      const char *gsrc = reinterpret_cast<const char *>(Gw) + lane * 16;
      const unsigned rdst = (unsigned)__cvta_generic_to_shared(ring) + lane * 16;
      auto fetch = [&](int s_) { // column of visit position s_ -> ring stage s_ % VC_RING (16-byte chunks)
        if (s_ < m) {
          const char *src = gsrc + soff[s_];
          const unsigned dst = rdst + (unsigned)(s_ % VC_RING) * (unsigned)(RS * 8);
          if (2 * lane < m0) vc_cp16<0>(dst, src);
          if ((NU + 1) / 2 > 1 && 2 * (lane + 32) < m0) vc_cp16<512>(dst, src);
          if ((NU + 1) / 2 > 2 && 2 * (lane + 64) < m0) vc_cp16<1024>(dst, src);
          if ((NU + 1) / 2 > 3 && 2 * (lane + 96) < m0) vc_cp16<1536>(dst, src);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      };
#pragma unroll
      for (int d = 0; d < VC_RING - 1; ++d) fetch(d);
      double maxH = 0.0;
#pragma unroll
      for (int u = 0; u < MS; ++u) {
        const int cnt = min(32, m - 32 * u);
        for (int i = 0; i < cnt; ++i) { // visit position s_ = 32 u + i, ring stage i % VC_RING (32 % VC_RING == 0)
          const int s_ = 32 * u + i;
          asm volatile("cp.async.wait_group %0;" ::"n"(VC_RING - 2) : "memory"); // column of position s_ has landed
          __syncwarp();
          const double *col = ring + (i % VC_RING) * RS;
          double gc[MS];
#pragma unroll
          for (int q = 0; q < MS; ++q) gc[q] = col[row[q]];
          fetch(s_ + VC_RING - 1); // into the stage of position s_-1, which every lane read before the barrier above
          const double v = __dsub_rn(be[u], __dmul_rn(Ax[u] + cc[u], ai[u]));
          const double nwl = cd_shrink(v, th[u]);
          const double h = __shfl_sync(0xffffffffu, nwl - be[u], i);
          if (lane == i) be[u] = nwl;
          // h == 0 adds an exact zero (the Gram entries are finite): no data-dependent branch on the chain
#pragma unroll
          for (int q = 0; q < MS; ++q) Ax[q] = __dadd_rn(Ax[q], __dmul_rn(gc[q], h));
          accepted += h != 0.0;
          maxH = fmax(maxH, fabs(h));
        }
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
#pragma unroll
      for (int u = 0; u < MS; ++u) {
        const int s_ = lane + 32 * u;
        if (s_ < m) {
          sAxE[row[u]] = Ax[u];
          sval[slist[s_]] = be[u];
        }
      }
      __syncwarp();
      return maxH;
    };

    DevStats st;
    st.passes = st.full_passes = st.visits = st.accepted = 0;
    st.maxH = 0.0;
    st.converged = 0;
    st.outer_iters = 0;
    st.sigma = 0.0;
    unsigned long long pass_counter = 0;
    long long pc[6] = {0, 0, 0, 0, 0, 0};
    const long long tstart = clock64();
    // _coordinateDescent! (coordinate_descent.jl:65-92) from the current iterate: full pass first, active-set passes
    // until one converges, exit when a full pass has max|h| < optTol
    auto run_solve = [&]() {
    bool conv = true;
    long long iter = 0;
    st.converged = 0;
    while (iter < a.maxIter) {
      double maxH = 0.0;
      iter += 1;
      st.passes += 1;
      long long tq = clock64();
      if (conv) { // ---- full pass: speculative first-mover scan (exact Gauss-Seidel order)
        if (in_phase) {
          phase_close(nact);
          pc[4] += clock64() - tq;
          tq = clock64();
        }
        st.full_passes += 1;
        st.visits += ep;
        const PermKey pk = cd_perm_key((uint32_t)ep, a.seed, pass_counter);
        const int m_old = nact;
        long long cur = -1;
        for (;;) {
          unsigned best = NONE;
          double bh = 0.0, bnw = 0.0;
          int bk = 0;
#pragma unroll
          for (int u = 0; u < NU; ++u) {
            const int t = lane + 32 * u;
            if (t < ep) {
              const unsigned key = ordered ? (unsigned)t : cd_perm_inv(pk, (uint32_t)t);
              if ((long long)key > cur) {
                const double tt = __dmul_rn(Ax[u] + cc[u], ai[u]);
                const double v = __dsub_rn(be[u], tt);
                const double nw = cd_shrink(v, th[u]);
                const double h = nw - be[u];
                if (h != 0.0 && key < best) {
                  best = key;
                  bh = h;
                  bnw = nw;
                  bk = t;
                }
              }
            }
          }
          const unsigned wmin = __reduce_min_sync(0xffffffffu, best);
          if (wmin == NONE) break;
          const int src = __ffs(__ballot_sync(0xffffffffu, best == wmin)) - 1;
          const int k = __shfl_sync(0xffffffffu, bk, src);
          const double h = __shfl_sync(0xffffffffu, bh, src), nw = __shfl_sync(0xffffffffu, bnw, src);
          if (lane == (k & 31)) {
#pragma unroll
            for (int u = 0; u < NU; ++u)
              if (u == (k >> 5)) be[u] = nw;
          }
          apply(k, h);
          if (!sin[k]) { // setindex! appends on the first non-zero store
            __syncwarp();
            if (lane == 0) {
              sin[k] = 1;
              sact[nact] = k;
            }
            nact += 1;
            __syncwarp();
          }
          maxH = fmax(maxH, fabs(h));
          st.accepted += 1;
          cur = (long long)wmin;
        }
        pc[0] += clock64() - tq;
        tq = clock64();
        // list order after the pass (common.cuh: cd_compact_list); the visited non-members whose tentative
        // value is exactly zero (not appended by the reference) are not tracked here: it needs an exactly
        // zero gradient and only affects the visit order of later active-set passes
#pragma unroll
        for (int u = 0; u < NU; ++u)
          if (lane + 32 * u < ep) sbeta[lane + 32 * u] = be[u];
        __syncwarp();
        for (int i = lane; i < nact; i += 32) sval[i] = sbeta[sact[i]];
        for (int e = m_old + lane; e < nact; e += 32) {
          const int k = sact[e];
          const int vis = ordered ? k : (int)cd_perm_inv(pk, (uint32_t)k);
          int before = 0;
          for (int j = 0; j < m_old; ++j) before += (ordered ? sact[j] : (int)cd_perm_inv(pk, (uint32_t)sact[j])) < vis;
          snewpos[e - m_old] = m_old + vis - before;
        }
        __syncwarp();
        cd_compact_list<32>(sact, sval, m_old, nact, snewpos, sin, scmp, stmpd, s2);
        nact = s2[0];
        __syncwarp();
        pc[2] += clock64() - tq;
      } else { // ---- active-set pass: sequential chain over the stored entries
        const int m = nact;
        st.visits += m;
        if (!in_phase) {
          phase_open(m);
          pc[3] += clock64() - tq;
          tq = clock64();
        }
        const PermKey pkm = cd_perm_key((uint32_t)max(m, 1), a.seed, pass_counter);
        int acc_pass = 0;
        {
          const int ms = (m + 31) >> 5;
          if (ms <= 1 || NU == 1)
            maxH = phase_pass(m, pkm, acc_pass, std::integral_constant<int, 1>{});
          else if (ms == 2 || NU == 2)
            maxH = phase_pass(m, pkm, acc_pass, std::integral_constant<int, (NU < 2 ? NU : 2)>{});
          else if (ms == 3 || NU == 3)
            maxH = phase_pass(m, pkm, acc_pass, std::integral_constant<int, (NU < 3 ? NU : 3)>{});
          else if (ms == 4 || NU == 4)
            maxH = phase_pass(m, pkm, acc_pass, std::integral_constant<int, (NU < 4 ? NU : 4)>{});
          else
            maxH = phase_pass(m, pkm, acc_pass, std::integral_constant<int, NU>{});
        }
        st.accepted += acc_pass;
        pc[1] += clock64() - tq;
        tq = clock64();
        cd_compact_list<32>(sact, sval, m, m, snewpos, sin, scmp, stmpd, s2); // dropzeros!
        nact = s2[0];
        __syncwarp();
        pc[2] += clock64() - tq;
      }
      pass_counter += 1;
      st.maxH = maxH;
      const bool prev = conv;
      conv = maxH < a.optTol;
      if (prev && conv) {
        st.converged = 1;
        break;
      }
    }
    if (in_phase) phase_close(nact); // pass budget ran out inside a phase: back to the coordinate layout
    };

    // A_g[S,S] x = rhs for the ms coordinates listed (ascending) in sact; rhs and the solution live in stmpd.
    // Left-looking Cholesky in this warp's scratch (columns contiguous: coalesced; all loads of a column's update
    // are independent), then the two triangular solves.  false: not positive definite.
    auto spd_solve = [&](int ms) -> bool {
      const int ldm = (ms + 1) & ~1;
      double *rhs = stmpd, *lrow = sval; // lrow: L[j, 0..j) of the column being formed
      bool ok = true;
      for (int j = 0; j < ms; ++j) {
        const int kj = sact[j];
        for (int k2 = lane; k2 < j; k2 += 32) lrow[k2] = Gw[j + (long long)k2 * ldm];
        __syncwarp();
        double acc[NU];
#pragma unroll
        for (int u = 0; u < NU; ++u) {
          const int i = j + lane + 32 * u;
          acc[u] = i < ms ? moment(sact[i], kj) : 0.0;
        }
        for (int k2 = 0; k2 < j; ++k2) {
          const double ljk = lrow[k2];
          const double *ck = Gw + (long long)k2 * ldm + j + lane;
#pragma unroll
          for (int u = 0; u < NU; ++u)
            if (j + lane + 32 * u < ms) acc[u] = fma(-__ldcg(ck + 32 * u), ljk, acc[u]);
        }
        const double djj = __shfl_sync(0xffffffffu, acc[0], 0);
        if (!(djj > 0.0)) ok = false;
        const double d = sqrt(djj);
#pragma unroll
        for (int u = 0; u < NU; ++u) {
          const int i = j + lane + 32 * u;
          if (i < ms) Gw[i + (long long)j * ldm] = i == j ? d : acc[u] / d;
        }
        __syncwarp();
      }
      for (int j = 0; j < ms; ++j) { // L y = rhs (column sweep)
        const double yj = rhs[j] / __ldcg(Gw + j + (long long)j * ldm);
        __syncwarp();
        if (lane == 0) rhs[j] = yj;
        for (int i = j + 1 + lane; i < ms; i += 32) rhs[i] = fma(-__ldcg(Gw + i + (long long)j * ldm), yj, rhs[i]);
        __syncwarp();
      }
      for (int j = ms - 1; j >= 0; --j) { // L' x = y (dot products)
        double sacc = 0.0;
        for (int i = j + 1 + lane; i < ms; i += 32) sacc = fma(__ldcg(Gw + i + (long long)j * ldm), rhs[i], sacc);
        sacc = warp_sum(sacc);
        const double xj = (rhs[j] - sacc) / __ldcg(Gw + j + (long long)j * ldm);
        __syncwarp();
        if (lane == 0) rhs[j] = xj;
        __syncwarp();
      }
      return ok;
    };
    // the expanded coordinates of every group with a non-zero coefficient, ascending, into sact
    // (get_nonzero_coordinates!(S, beta, p, degree, true), varying_coefficient_lasso.jl:488-512); returns |S|
    auto selected_groups = [&]() -> int {
      for (int k2 = lane; k2 < ep; k2 += 32) sin[k2] = 0;
      __syncwarp();
#pragma unroll
      for (int u = 0; u < NU; ++u)
        if (lane + 32 * u < ep && be[u] != 0.0) {
          const int j0 = tj[u] * dg;
          for (int l = 0; l < dg; ++l) sin[j0 + l] = 1; // the whole group (same byte value from every writer)
        }
      __syncwarp();
      int ms = 0;
      for (int k0 = 0; k0 < ep; k0 += 32) {
        const int k2 = k0 + lane;
        const bool inS = k2 < ep && sin[k2] != 0;
        const unsigned bal = __ballot_sync(0xffffffffu, inS);
        if (inS) sact[ms + __popc(bal & ((1u << lane) - 1u))] = k2;
        ms += __popc(bal);
      }
      __syncwarp();
      return ms;
    };

    if constexpr (!LVO) {
      run_solve();
    } else {
      // ---- one problem of lvocv_locpolyl1 (varying_coefficient_lasso.jl:81-137): scaled-lasso sigma loop on the
      // leave-one-out local problem, refit, prediction of the left-out response.  Everything about the residual
      // comes from the moment blocks: r'Wr/n = y'Wy/n + 2 b'beta + beta'(A beta), sum(w)/n.
      const double yWy = __ldg(Cg + a.P2 + a.p), Sw = __ldg(Cg + a.P2 + a.p + 1);
      auto sigma_now = [&]() -> double { // _getSigma(w, f.r), utils.jl:167-175
        double t = 0.0;
#pragma unroll
        for (int u = 0; u < NU; ++u) t += be[u] * (2.0 * cc[u] + Ax[u]);
        t = warp_sum(t);
        return sqrt(fmax(yWy + t, 0.0) / Sw);
      };
      // _findInitResiduals!(w, wX, y, min(10, ep), f.r) (utils.jl:79-92): the s columns with the largest |X_k'Wy|,
      // weighted least squares on them, sigma from those residuals
      const int s_init = min(10, ep);
      double thr = 0.0;
      {
        double cv[NU];
#pragma unroll
        for (int u = 0; u < NU; ++u) cv[u] = lane + 32 * u < ep ? fabs(cc[u]) : -1.0;
        for (int r = 0; r < s_init; ++r) { // r-th largest by repeated arg-max
          double best = -1.0;
#pragma unroll
          for (int u = 0; u < NU; ++u) best = fmax(best, cv[u]);
          const double wbest = warp_max(best);
          thr = wbest;
          const unsigned bal = __ballot_sync(0xffffffffu, best == wbest);
          if (lane == __ffs(bal) - 1) {
            bool done = false;
#pragma unroll
            for (int u = 0; u < NU; ++u)
              if (!done && cv[u] == wbest) {
                cv[u] = -1.0;
                done = true;
              }
          }
        }
      }
      int ms0 = 0;
      for (int k0 = 0; k0 < ep; k0 += 32) { // S = storage .>= nlargest(s, storage)[end] (ties included)
        const int k2 = k0 + lane;
        const bool inS = k2 < ep && fabs(scc[k2]) >= thr;
        const unsigned bal = __ballot_sync(0xffffffffu, inS);
        if (inS) sact[ms0 + __popc(bal & ((1u << lane) - 1u))] = k2;
        ms0 += __popc(bal);
      }
      __syncwarp();
      for (int i = lane; i < ms0; i += 32) stmpd[i] = -scc[sact[i]];
      __syncwarp();
      spd_solve(ms0);
      double t0 = 0.0;
      for (int i = lane; i < ms0; i += 32) t0 = fma(scc[sact[i]], stmpd[i], t0); // at the LS solution r'Wr/n = y'Wy/n + b_S'gamma
      t0 = warp_sum(t0);
      double sigma = sqrt(fmax(yWy + t0, 0.0) / Sw);
      __syncwarp();
      for (int outer = 1; outer <= 10; ++outer) { // :115-124
        const double lam = a.lambda0 * sigma;
#pragma unroll
        for (int u = 0; u < NU; ++u)
          if (lane + 32 * u < ep) {
            th[u] = __dmul_rn(__dmul_rn(ai[u], lam), sqrt(1.0 / ai[u])); // lambda0 sigma omega_k / A_kk
            sth[lane + 32 * u] = th[u];
          }
        __syncwarp();
        run_solve();
        st.outer_iters = outer;
        const double snew = sigma_now();
        if (fabs(snew - sigma) / sigma < 1e-2) break;
        sigma = snew;
      }
      st.sigma = sigma;
    }
    if (a.prof && lane == 0) {
      pc[5] = clock64() - tstart;
      for (int i = 0; i < 6; ++i) atomicAdd(a.prof + i, (unsigned long long)pc[i]);
      atomicMax(a.prof + 6, (unsigned long long)pc[5]);
    }
    if (a.out) {
      double *col = a.out + (long long)g * ep;
#pragma unroll
      for (int u = 0; u < NU; ++u)
        if (lane + 32 * u < ep) col[lane + 32 * u] = be[u];
    }
    if (lane == 0 && a.stats) a.stats[g] = st;
    __syncwarp();
    if (LVO || a.outR) {
      // ---- refit (varying_coefficient_lasso.jl:71-76): A_g[S,S] x = -b_g[S] on the expanded coordinates S of every
      // group with a non-zero coefficient; both sides come from the moment blocks
      const int ms = selected_groups();
      for (int i = lane; i < ms; i += 32) stmpd[i] = -scc[sact[i]];
      __syncwarp();
      const bool ok = ms > 0 ? spd_solve(ms) : true;
      if (a.outR) {
        double *colR = a.outR + (long long)g * ep;
        for (int k2 = lane; k2 < ep; k2 += 32) colR[k2] = 0.0;
        __syncwarp();
        for (int i = lane; i < ms; i += 32) colR[sact[i]] = ok ? stmpd[i] : nan("");
      }
      if constexpr (LVO) {
        // prediction of the left-out response (:129-131): z0 = z_i, so only the degree-0 columns of row i are
        // non-zero: Yh = sum_{(j,0) in S} X[i,j] x_(j,0)
        const int i_obs = g % a.n;
        double yh = 0.0;
        for (int i = lane; i < ms; i += 32) {
          const int k2 = sact[i], j = k2 / dg;
          if (k2 - j * dg == 0) yh = fma(__ldg(a.X + i_obs + (long long)j * a.ldx), stmpd[i], yh);
        }
        yh = warp_sum(yh);
        if (lane == 0) {
          const double e = yh - __ldg(a.y + i_obs);
          a.lvo_err[g] = ok ? e * e : nan("");
        }
      }
      __syncwarp();
    }
  }
}

} // namespace

// host driver of the moment form; grid points are processed in chunks so the moment blocks stay <= ~1 GiB
static int vc_solve_moment(const double *X, int64_t n, int64_t p, int64_t ldx, const double *z, const double *y,
                           const double *zgrid, int64_t m, int64_t m_begin, int64_t m_end, int degree, int kernel_kind,
                           double bandwidth, double lambda0, const cdgpu_options *opt, int device, double *out,
                           double *outR, cdgpu_stats *stats, const double *harr = nullptr, double *lvo_err = nullptr) {
  // harr != null: the problems are those of lvocv_locpolyl1 (m = numH * n, zgrid unused, out/outR null, one squared
  // prediction error per problem into lvo_err)
  const int64_t ep = p * (degree + 1), mloc = m_end - m_begin, P2 = p * (p + 1) / 2, PA = P2 + p + 2;
  const int nq = 2 * degree + 1;
  const long long ldz = (n + 1) & ~(int64_t)1, ldc = (PA + 1) & ~(int64_t)1;
  int64_t chunk = (int64_t)((1ll << 30) / ((long long)ldc * nq * 8));
  chunk = std::max<int64_t>(1, std::min<int64_t>(chunk, mloc));
  chunk = std::min<int64_t>(chunk, std::max<int64_t>(1, (int64_t)((1ll << 30) / ((long long)ldz * nq * 8))));
  if (const char *env = getenv("CDGPU_VC_CHUNK")) chunk = std::max<int64_t>(1, std::min<int64_t>(chunk, atoll(env))); // tests: force several chunks
  cudaStream_t s = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr, eg = nullptr;
  double *dX = nullptr, *dz = nullptr, *dy = nullptr, *dgz = nullptr, *dout = nullptr, *dZ = nullptr, *dV = nullptr, *dC = nullptr;
  DevStats *dst = nullptr;
  double *dG = nullptr, *doutR = nullptr, *dharr = nullptr, *derr = nullptr;
  int *dcounter = nullptr;
  unsigned long long *dprof = nullptr;
  void *dtiles = nullptr;
  int rc = CDGPU_OK;
  auto cleanup = [&]() {
    if (s) cudaStreamSynchronize(s);
    for (void *ptr : {(void *)dX, (void *)dz, (void *)dy, (void *)dgz, (void *)dout, (void *)dZ, (void *)dV, (void *)dC,
                      (void *)dst, (void *)dcounter, (void *)dprof, (void *)dG, (void *)doutR, (void *)dharr, (void *)derr, dtiles})
      if (ptr) cudaFreeAsync(ptr, s);
    if (eg) cudaEventDestroy(eg);
    if (s) {
      cudaStreamSynchronize(s);
      stream_set_release(device, StreamSet{s, e0, e1});
    }
  };
#define VM_TRY(expr)                                                                                        \
  do {                                                                                                      \
    cudaError_t _e = (expr);                                                                                \
    if (_e != cudaSuccess) {                                                                                \
      rc = cdgpu_set_error(_e == cudaErrorMemoryAllocation ? CDGPU_ENOMEM : CDGPU_ECUDA, "%s: %s", #expr,   \
                           cudaGetErrorString(_e));                                                         \
      cleanup();                                                                                            \
      return rc;                                                                                            \
    }                                                                                                       \
  } while (0)
  {
    StreamSet ss;
    rc = stream_set_acquire(device, &ss); // recycled: creating streams/events goes through the resource manager
    if (rc) return rc;
    s = ss.stream;
    e0 = ss.ev0;
    e1 = ss.ev1;
  }
  VM_TRY(cudaEventCreateWithFlags(&eg, cudaEventDefault));
  VM_TRY(cudaMallocAsync((void **)&dX, (size_t)n * p * sizeof(double), s));
  VM_TRY(cudaMallocAsync((void **)&dz, (size_t)n * sizeof(double), s));
  VM_TRY(cudaMallocAsync((void **)&dy, (size_t)n * sizeof(double), s));
  if (!harr) VM_TRY(cudaMallocAsync((void **)&dgz, (size_t)m * sizeof(double), s));
  if (out) VM_TRY(cudaMallocAsync((void **)&dout, (size_t)ep * m * sizeof(double), s));
  if (harr) {
    const int64_t numH = m / n;
    VM_TRY(cudaMallocAsync((void **)&dharr, (size_t)numH * sizeof(double), s));
    VM_TRY(cudaMallocAsync((void **)&derr, (size_t)m * sizeof(double), s));
    VM_TRY(cudaMemcpyAsync(dharr, harr, (size_t)numH * sizeof(double), cudaMemcpyHostToDevice, s));
  }
  VM_TRY(cudaMallocAsync((void **)&dst, (size_t)m * sizeof(DevStats), s));
  if (outR) VM_TRY(cudaMallocAsync((void **)&doutR, (size_t)ep * m * sizeof(double), s));
  VM_TRY(cudaMallocAsync((void **)&dZ, (size_t)ldz * PA * sizeof(double), s));
  VM_TRY(cudaMallocAsync((void **)&dV, (size_t)ldz * nq * chunk * sizeof(double), s));
  VM_TRY(cudaMallocAsync((void **)&dC, (size_t)ldc * nq * chunk * sizeof(double), s));
  VM_TRY(cudaMallocAsync((void **)&dcounter, sizeof(int), s));
  if (getenv("CDGPU_PROFILE")) {
    VM_TRY(cudaMallocAsync((void **)&dprof, 8 * sizeof(unsigned long long), s));
    VM_TRY(cudaMemsetAsync(dprof, 0, 8 * sizeof(unsigned long long), s));
  }
  VM_TRY(cudaMemcpy2DAsync(dX, n * sizeof(double), X, ldx * sizeof(double), n * sizeof(double), p, cudaMemcpyHostToDevice, s));
  VM_TRY(cudaMemcpyAsync(dz, z, n * sizeof(double), cudaMemcpyHostToDevice, s));
  VM_TRY(cudaMemcpyAsync(dy, y, n * sizeof(double), cudaMemcpyHostToDevice, s));
  if (!harr) VM_TRY(cudaMemcpyAsync(dgz, zgrid, m * sizeof(double), cudaMemcpyHostToDevice, s));
  int sms = 0;
  VM_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  const int nu = (int)((ep + 31) / 32);
  const void *kfn_std = nu <= 1   ? (const void *)vc_cov_kernel<1, false>
                        : nu == 2 ? (const void *)vc_cov_kernel<2, false>
                        : nu == 3 ? (const void *)vc_cov_kernel<3, false>
                        : nu == 4 ? (const void *)vc_cov_kernel<4, false>
                        : nu == 5 ? (const void *)vc_cov_kernel<5, false>
                        : nu == 6 ? (const void *)vc_cov_kernel<6, false>
                                  : (const void *)vc_cov_kernel<8, false>;
  const void *kfn_lvo = nu <= 1   ? (const void *)vc_cov_kernel<1, true>
                        : nu == 2 ? (const void *)vc_cov_kernel<2, true>
                        : nu == 3 ? (const void *)vc_cov_kernel<3, true>
                        : nu == 4 ? (const void *)vc_cov_kernel<4, true>
                        : nu == 5 ? (const void *)vc_cov_kernel<5, true>
                        : nu == 6 ? (const void *)vc_cov_kernel<6, true>
                                  : (const void *)vc_cov_kernel<8, true>;
  const void *kfn = harr ? kfn_lvo : kfn_std;
  const size_t wsz = vc_cov_warp_bytes((int)ep, nu <= 6 ? std::max(nu, 1) : 8);
  const size_t dyn = VCW * wsz;
  VM_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  int occ = 0;
  VM_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, VCW * 32, dyn));
  if (occ < 1) occ = 1;
  const int64_t max_ctas = std::min<int64_t>((std::min<int64_t>(chunk, mloc) + VCW - 1) / VCW, (int64_t)occ * sms);
  const int64_t MCs = (ep + 1) & ~(int64_t)1;
  VM_TRY(cudaMallocAsync((void **)&dG, (size_t)max_ctas * VCW * MCs * MCs * sizeof(double), s));
  VM_TRY(cudaEventRecord(e0, s));
  vc_build_z_kernel<<<(unsigned)PA, 128, 0, s>>>(dX, n, (int)n, (int)p, dy, dZ, ldz);
  VM_TRY(cudaGetLastError());
  CD_COUNT_LAUNCH(1);
  for (int64_t c0 = m_begin; c0 < m_end; c0 += chunk) {
    const int64_t c1 = std::min<int64_t>(m_end, c0 + chunk), mc = c1 - c0;
    vc_build_v_kernel<<<(unsigned)mc, 128, 0, s>>>(dz, dgz, (int)n, (int)c0, (int)c1, nq, kernel_kind, bandwidth, dV, ldz, dharr);
    VM_TRY(cudaGetLastError());
    if (dtiles) {
      VM_TRY(cudaFreeAsync(dtiles, s));
      dtiles = nullptr;
    }
    rc = launch_gemm_tn(s, sms, dZ, (int)PA, ldz, dV, (int)(mc * nq), ldz, n, dC, ldc, (double)n, &dtiles);
    if (rc) {
      cleanup();
      return rc;
    }
    VM_TRY(cudaEventRecord(eg, s)); // (last chunk's) moment blocks formed
    VM_TRY(cudaMemsetAsync(dcounter, 0, sizeof(int), s));
    VcCovArgs a = {};
    a.C = dC;
    a.ldc = ldc;
    a.p = (int)p;
    a.degree = degree;
    a.ep = (int)ep;
    a.P2 = (int)P2;
    a.g0 = (int)c0;
    a.g1 = (int)c1;
    a.lambda0 = lambda0;
    a.maxIter = opt->maxIter;
    a.optTol = opt->optTol;
    a.randomize = opt->randomize;
    a.seed = opt->seed;
    a.out = dout;
    a.outR = doutR;
    a.stats = dst;
    a.counter = dcounter;
    a.lvo = harr ? 1 : 0;
    a.n = (int)n;
    a.X = dX;
    a.y = dy;
    a.ldx = n;
    a.lvo_err = derr;
    a.prof = dprof;
    a.gscr = dG;
    const int64_t ctas = std::min<int64_t>((mc + VCW - 1) / VCW, (int64_t)occ * sms);
    void *kargs[] = {(void *)&a};
    VM_TRY(cudaLaunchKernel(kfn, dim3((unsigned)ctas), dim3(VCW * 32), kargs, dyn, s));
    CD_COUNT_LAUNCH(2);
  }
  VM_TRY(cudaEventRecord(e1, s));
  if (out)
    VM_TRY(cudaMemcpyAsync(out + m_begin * ep, dout + m_begin * ep, (size_t)mloc * ep * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (lvo_err)
    VM_TRY(cudaMemcpyAsync(lvo_err + m_begin, derr + m_begin, (size_t)mloc * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (outR)
    VM_TRY(cudaMemcpyAsync(outR + m_begin * ep, doutR + m_begin * ep, (size_t)mloc * ep * sizeof(double), cudaMemcpyDeviceToHost, s));
  VM_TRY(cudaStreamSynchronize(s));
  if (stats) {
    float ms = 0.f;
    VM_TRY(cudaEventElapsedTime(&ms, e0, e1));
    if (getenv("CDGPU_PROFILE")) {
      float mg = 0.f;
      cudaEventElapsedTime(&mg, e0, eg);
      fprintf(stderr, "[cdgpu profile] vc moment form: total %.3f ms, of which Z/V build + DMMA GEMM %.3f ms (single chunk: %s)\n", ms, mg,
              chunk >= mloc ? "yes" : "no, GEMM time is the last chunk's offset");
      if (dprof) {
        unsigned long long pf[8];
        cudaMemcpy(pf, dprof, sizeof pf, cudaMemcpyDeviceToHost);
        fprintf(stderr, "[cdgpu profile]   warp cycles summed over problems (M): full passes %.1f | active chain %.1f | compaction %.1f | phase open %.1f | "
                        "phase close %.1f | total %.1f | longest problem %.2f\n",
                pf[0] * 1e-6, pf[1] * 1e-6, pf[2] * 1e-6, pf[3] * 1e-6, pf[4] * 1e-6, pf[5] * 1e-6, pf[6] * 1e-6);
      }
    }
    std::vector<DevStats> hst((size_t)mloc);
    VM_TRY(cudaMemcpy(hst.data(), dst + m_begin, (size_t)mloc * sizeof(DevStats), cudaMemcpyDeviceToHost));
    for (int64_t g = 0; g < mloc; ++g) {
      cdgpu_stats *o = stats + m_begin + g;
      o->passes = hst[g].passes;
      o->full_passes = hst[g].full_passes;
      o->visits = hst[g].visits;
      o->accepted = hst[g].accepted;
      o->maxH = hst[g].maxH;
      o->converged = hst[g].converged;
      o->outer_iters = hst[g].outer_iters;
      o->sigma = hst[g].sigma;
      o->device_ms = g == 0 ? (double)ms : 0.0;
    }
  }
#undef VM_TRY
  cleanup();
  return CDGPU_OK;
}

static int vc_solve_impl(const double *X, int64_t n, int64_t p, int64_t ldx, const double *z, const double *y,
                         const double *zgrid, int64_t m, int64_t m_begin, int64_t m_end, int degree, int kernel_kind,
                         double bandwidth, double lambda0, const cdgpu_options *opt, int device, double *out, double *outR,
                         cdgpu_stats *stats);
// lvocv_locpolyl1 (varying_coefficient_lasso.jl:81-137): numH * n leave-one-out local scaled-lasso problems, all in
// one batch; problems [q_begin, q_end) of the (bandwidth-major) list are solved (sharding hook), MSE[h] is summed on
// the host in observation order from the per-problem squared errors.
API int cdgpu_vc_lvocv(const double *X, int64_t n, int64_t p, int64_t ldx, const double *z, const double *y, int degree,
                       const double *hArr, int64_t numH, int kernel_kind, double lambda0, const cdgpu_options *opt,
                       int64_t q_begin, int64_t q_end, int device, double *sqerr, cdgpu_stats *stats) {
  return api_guard([&]() -> int {
  if (!X || !z || !y || !hArr || !opt || !sqerr) return cdgpu_set_error(CDGPU_EARG, "null pointer");
  const int64_t m = numH * n;
  if (n < 2 || p < 1 || ldx < n || degree < 0 || numH < 0 || q_begin < 0 || q_end > m || q_begin > q_end)
    return cdgpu_set_error(CDGPU_EDIM, "DimensionMismatch");
  if (kernel_kind != CDGPU_KERNEL_GAUSSIAN && kernel_kind != CDGPU_KERNEL_EPANECHNIKOV)
    return cdgpu_set_error(CDGPU_EARG, "unknown smoothing kernel");
  if (opt->maxIter < 0 || opt->randomize < 0 || opt->randomize > 1) return cdgpu_set_error(CDGPU_EARG, "bad options");
  const int64_t ep = p * (degree + 1);
  if (n > 0x7fffffff || ep > 0x7fffffff || m > 0x7fffffff) return cdgpu_set_error(CDGPU_EDIM, "sizes must fit in 31 bits");
  if (ep > 256) return cdgpu_set_error(CDGPU_ECAP, "lvocv needs the moment form: p*(degree+1) = %lld exceeds 256", (long long)ep);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return cdgpu_set_error(CDGPU_ENODEV, "no CUDA device (%s); libcdgpu has no CPU fallback",
                           e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return cdgpu_set_error(CDGPU_EARG, "device out of range");
  CD_TRY(cd_use_device(device));
  if (q_begin == q_end) return CDGPU_OK;
  return vc_solve_moment(X, n, p, ldx, z, y, nullptr, m, q_begin, q_end, degree, kernel_kind, 0.0, lambda0, opt, device, nullptr,
                         nullptr, stats, hArr, sqerr);
  });
}
API int cdgpu_vc_solve(const double *X, int64_t n, int64_t p, int64_t ldx, const double *z, const double *y,
                       const double *zgrid, int64_t m, int64_t m_begin, int64_t m_end, int degree, int kernel_kind,
                       double bandwidth, double lambda0, const cdgpu_options *opt, int device, double *out,
                       cdgpu_stats *stats) {
  return api_guard([&]() -> int {
  return vc_solve_impl(X, n, p, ldx, z, y, zgrid, m, m_begin, m_end, degree, kernel_kind, bandwidth, lambda0, opt, device, out,
                       nullptr, stats);
  });
}
API int cdgpu_vc_solve_refit(const double *X, int64_t n, int64_t p, int64_t ldx, const double *z, const double *y,
                             const double *zgrid, int64_t m, int64_t m_begin, int64_t m_end, int degree, int kernel_kind,
                             double bandwidth, double lambda0, const cdgpu_options *opt, int device, double *out,
                             double *outR, cdgpu_stats *stats) {
  return api_guard([&]() -> int {
  if (!outR) return cdgpu_set_error(CDGPU_EARG, "null pointer");
  return vc_solve_impl(X, n, p, ldx, z, y, zgrid, m, m_begin, m_end, degree, kernel_kind, bandwidth, lambda0, opt, device, out,
                       outR, stats);
  });
}
static int vc_solve_impl(const double *X, int64_t n, int64_t p, int64_t ldx, const double *z, const double *y,
                         const double *zgrid, int64_t m, int64_t m_begin, int64_t m_end, int degree, int kernel_kind,
                         double bandwidth, double lambda0, const cdgpu_options *opt, int device, double *out, double *outR,
                         cdgpu_stats *stats) {
  if (!X || !z || !y || !zgrid || !opt || !out) return cdgpu_set_error(CDGPU_EARG, "null pointer");
  if (n < 1 || p < 1 || ldx < n || degree < 0 || m < 0 || m_begin < 0 || m_end > m || m_begin > m_end)
    return cdgpu_set_error(CDGPU_EDIM, "DimensionMismatch");
  if (kernel_kind != CDGPU_KERNEL_GAUSSIAN && kernel_kind != CDGPU_KERNEL_EPANECHNIKOV)
    return cdgpu_set_error(CDGPU_EARG, "unknown smoothing kernel");
  if (opt->maxIter < 0 || opt->randomize < 0 || opt->randomize > 1) return cdgpu_set_error(CDGPU_EARG, "bad options");
  const int64_t ep = p * (degree + 1);
  if (n > 0x7fffffff || ep > 0x7fffffff || m > 0x7fffffff) return cdgpu_set_error(CDGPU_EDIM, "sizes must fit in 31 bits");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return cdgpu_set_error(CDGPU_ENODEV, "no CUDA device (%s); libcdgpu has no CPU fallback",
                           e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return cdgpu_set_error(CDGPU_EARG, "device out of range");
  CD_TRY(cd_use_device(device));
  const int64_t mloc = m_end - m_begin;
  if (mloc == 0) return CDGPU_OK;
  {
    // moment (covariance) form unless the expanded problem is too wide for the warp kernel's registers;
    // CDGPU_VC_FORM=naive keeps the residual-form kernels below
    const char *form = getenv("CDGPU_VC_FORM");
    if (ep <= 256 && (outR || !(form && strcmp(form, "naive") == 0)))
      return vc_solve_moment(X, n, p, ldx, z, y, zgrid, m, m_begin, m_end, degree, kernel_kind, bandwidth, lambda0, opt,
                             device, out, outR, stats);
    if (outR)
      return cdgpu_set_error(CDGPU_ECAP, "refit on the device needs the moment form: p*(degree+1) = %lld exceeds 256",
                             (long long)ep);
  }
  int nr = n <= 128 ? 4 : (n <= 256 ? 8 : (n <= 512 ? 16 : 0)); // 0: CTA-per-problem kernel
  if (const char *env = getenv("CDGPU_VC_THREADS")) nr = atoi(env) == 32 ? nr : 0;
  const size_t dyn_cta = (sizeof(VSm) + 15) / 16 * 16 + (size_t)(3 * n + 4 * ep) * sizeof(double) + (size_t)ep * (8 * 4 + 1) + 16;
  const size_t dyn_warp = (size_t)4 * ep * sizeof(double) + (size_t)(8 * ep + 4) * sizeof(int) + (size_t)ep + 16;
  if (nr && dyn_warp > 227 * 1024) nr = 0;
  const size_t dyn = nr ? dyn_warp : dyn_cta;
  if (dyn > 227 * 1024)
    return cdgpu_set_error(CDGPU_ECAP, "local problem does not fit in shared memory (n=%lld, ep=%lld)", (long long)n,
                           (long long)ep);
  double *dX = nullptr, *dz = nullptr, *dy = nullptr, *dgz = nullptr, *dout = nullptr;
  DevStats *dst = nullptr;
  cudaStream_t s = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  int rc = CDGPU_OK;
  auto cleanup = [&]() {
    cudaFree(dX);
    cudaFree(dz);
    cudaFree(dy);
    cudaFree(dgz);
    cudaFree(dout);
    cudaFree(dst);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (s) cudaStreamDestroy(s);
  };
#define VC_TRY(expr)                                                                                        \
  do {                                                                                                      \
    cudaError_t _e = (expr);                                                                                \
    if (_e != cudaSuccess) {                                                                                \
      rc = cdgpu_set_error(_e == cudaErrorMemoryAllocation ? CDGPU_ENOMEM : CDGPU_ECUDA, "%s: %s", #expr,   \
                           cudaGetErrorString(_e));                                                         \
      cleanup();                                                                                            \
      return rc;                                                                                            \
    }                                                                                                       \
  } while (0)
  VC_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  VC_TRY(cudaEventCreate(&e0));
  VC_TRY(cudaEventCreate(&e1));
  VC_TRY(cudaMalloc((void **)&dX, (size_t)n * p * sizeof(double)));
  VC_TRY(cudaMalloc((void **)&dz, (size_t)n * sizeof(double)));
  VC_TRY(cudaMalloc((void **)&dy, (size_t)n * sizeof(double)));
  VC_TRY(cudaMalloc((void **)&dgz, (size_t)m * sizeof(double)));
  VC_TRY(cudaMalloc((void **)&dout, (size_t)ep * m * sizeof(double)));
  VC_TRY(cudaMalloc((void **)&dst, (size_t)m * sizeof(DevStats)));
  VC_TRY(cudaMemcpy2DAsync(dX, n * sizeof(double), X, ldx * sizeof(double), n * sizeof(double), p,
                           cudaMemcpyHostToDevice, s));
  VC_TRY(cudaMemcpyAsync(dz, z, n * sizeof(double), cudaMemcpyHostToDevice, s));
  VC_TRY(cudaMemcpyAsync(dy, y, n * sizeof(double), cudaMemcpyHostToDevice, s));
  VC_TRY(cudaMemcpyAsync(dgz, zgrid, m * sizeof(double), cudaMemcpyHostToDevice, s));
  VcArgs a = {};
  a.X = dX;
  a.ldx = n;
  a.n = (int)n;
  a.p = (int)p;
  a.degree = degree;
  a.z = dz;
  a.y = dy;
  a.zgrid = dgz;
  a.g0 = (int)m_begin;
  a.g1 = (int)m_end;
  a.kernel_kind = kernel_kind;
  a.bandwidth = bandwidth;
  a.lambda0 = lambda0;
  a.maxIter = opt->maxIter;
  a.optTol = opt->optTol;
  a.randomize = opt->randomize;
  a.seed = opt->seed;
  a.out = dout;
  a.stats = dst;
  int VC_T = nr ? 32 : (n <= 1024 ? 64 : (n <= 4096 ? 128 : 256));
  if (const char *env = getenv("CDGPU_VC_THREADS")) {
    int v = atoi(env);
    if (!nr && (v == 64 || v == 128 || v == 256)) VC_T = v;
  }
  const void *kfn = nr == 4    ? (const void *)vc_warp_kernel<4>
                    : nr == 8  ? (const void *)vc_warp_kernel<8>
                    : nr == 16 ? (const void *)vc_warp_kernel<16>
                    : VC_T == 64 ? (const void *)vc_kernel<64>
                    : VC_T == 128 ? (const void *)vc_kernel<128> : (const void *)vc_kernel<256>;
  VC_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  int occ = 0, sms = 0;
  VC_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, VC_T, dyn));
  VC_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  if (occ < 1) occ = 1;
  const int grid = (int)(mloc < (int64_t)occ * sms ? mloc : (int64_t)occ * sms);
  VC_TRY(cudaEventRecord(e0, s));
  {
    void *kargs[] = {(void *)&a};
    VC_TRY(cudaLaunchKernel(kfn, dim3(grid), dim3(VC_T), kargs, dyn, s));
  }
  CD_COUNT_LAUNCH(1);
  VC_TRY(cudaGetLastError());
  VC_TRY(cudaEventRecord(e1, s));
  VC_TRY(cudaMemcpyAsync(out + m_begin * ep, dout + m_begin * ep, (size_t)mloc * ep * sizeof(double),
                         cudaMemcpyDeviceToHost, s));
  VC_TRY(cudaStreamSynchronize(s));
  if (stats) {
    float ms = 0.f;
    VC_TRY(cudaEventElapsedTime(&ms, e0, e1));
    DevStats *hst = new DevStats[(size_t)mloc];
    cudaError_t e2 = cudaMemcpy(hst, dst + m_begin, (size_t)mloc * sizeof(DevStats), cudaMemcpyDeviceToHost);
    if (e2 == cudaSuccess)
      for (int64_t g = 0; g < mloc; ++g) {
        cdgpu_stats *o = stats + m_begin + g;
        o->passes = hst[g].passes;
        o->full_passes = hst[g].full_passes;
        o->visits = hst[g].visits;
        o->accepted = hst[g].accepted;
        o->maxH = hst[g].maxH;
        o->converged = hst[g].converged;
        o->outer_iters = 0;
        o->sigma = 0.0;
        o->device_ms = g == 0 ? (double)ms : 0.0;
      }
    delete[] hst;
    VC_TRY(e2);
  }
#undef VC_TRY
  cleanup();
  return CDGPU_OK;
}
