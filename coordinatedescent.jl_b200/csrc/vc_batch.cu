// vc_batch.cu — K4: batched varying-coefficient lasso, locpolyl1 with refit=false
// (src/varying_coefficient_lasso.jl:30-79): for every grid point z0 one kernel-weighted
// local-polynomial lasso  min sum_i w_i (y_i - eX_i'b)^2/(2n) + lambda0 sum_k sd_k |b_k|,
//   w_i = K_h(z_i, z0) (:17-21),  eX[i,(j,l)] = X[i,j] (z_i - z0)^l (:550-569),
//   sd_k = sqrt(sum_i w_i eX_ik^2 / n) (utils.jl:140-151),
// solved by the reference's CD loop on CDWeightedLSLoss (cd_differentiable_function.jl:165-194).
//
// B200 design.  DEFAULT (ep <= 256; up to 512 for refit / chains / lvocv): the MOMENT FORM further down — all local Gram matrices of all grid points are
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
#include "vc_cov_kernel.cuh"

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

// ---- locpolyl1's SparseMatrixCSC result formed on the device (varying_coefficient_lasso.jl:46-47 spzeros(ep, m), :69
// out[:, i] = beta, :76): one warp per grid point counts, then writes, the stored entries of its column in row order
__global__ void vc_csc_count_kernel(const double *__restrict__ out, int ep, int g0, int g1, long long *cnt) {
  const int g = g0 + (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (g >= g1) return;
  const double *col = out + (long long)g * ep;
  int c = 0;
  for (int i = lane; i < ep; i += 32) c += col[i] != 0.0;
  for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) cnt[g - g0] = c;
}
__global__ void vc_csc_fill_kernel(const double *__restrict__ out, int ep, int g0, int g1, const long long *__restrict__ off,
                                   long long *rowval, double *nzval) {
  const int g = g0 + (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (g >= g1) return;
  const double *col = out + (long long)g * ep;
  long long o = off[g - g0];
  for (int i0 = 0; i0 < ep; i0 += 32) {
    const int i = i0 + lane;
    const double v = i < ep ? col[i] : 0.0;
    const unsigned bal = __ballot_sync(0xffffffffu, v != 0.0);
    if (v != 0.0) {
      const long long at = o + __popc(bal & ((1u << lane) - 1u));
      rowval[at] = i + 1;
      nzval[at] = v;
    }
    o += __popc(bal);
  }
}

} // namespace

// caller-side description of the CSC outputs of cdgpu_vc_solve_csc (R: the refit matrix, colptrR null without refit)
struct VcCsc {
  int64_t capacity;
  int64_t *colptr, *rowval;
  double *nzval;
  int64_t *colptrR, *rowvalR;
  double *nzvalR;
};
// dense ep x m result on the device -> CSC in the caller's arrays; only colptr and the stored entries are copied to the host.
// Columns outside [m_begin, m_end) are empty.  CDGPU_ECAP (colptr[m] = entries needed) when capacity is too small.
static int vc_emit_csc(cudaStream_t s, const double *dout, int64_t ep, int64_t m, int64_t m_begin, int64_t m_end, int64_t capacity,
                       int64_t *colptr, int64_t *rowval, double *nzval) {
  const int64_t mloc = m_end - m_begin;
  long long *dcnt = nullptr, *drow = nullptr;
  double *dval = nullptr;
  int rc = CDGPU_OK;
  auto fail = [&](cudaError_t e, const char *what) {
    rc = cdgpu_set_error(e == cudaErrorMemoryAllocation ? CDGPU_ENOMEM : CDGPU_ECUDA, "%s: %s", what, cudaGetErrorString(e));
  };
  std::vector<long long> cnt((size_t)mloc);
  cudaError_t e = cudaMallocAsync((void **)&dcnt, (size_t)mloc * sizeof(long long), s);
  const unsigned blocks = (unsigned)((mloc * 32 + 255) / 256);
  if (e != cudaSuccess) fail(e, "cudaMallocAsync(csc counts)");
  if (!rc) {
    vc_csc_count_kernel<<<blocks, 256, 0, s>>>(dout, (int)ep, (int)m_begin, (int)m_end, dcnt);
    CD_COUNT_LAUNCH(1);
    e = cudaMemcpyAsync(cnt.data(), dcnt, (size_t)mloc * sizeof(long long), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) fail(e, "csc count");
  }
  int64_t tot = 0;
  if (!rc) {
    for (int64_t g = 0; g <= m_begin; ++g) colptr[g] = 0;
    for (int64_t g = 0; g < mloc; ++g) {
      const long long c = cnt[(size_t)g];
      cnt[(size_t)g] = tot; // exclusive offsets for the fill
      tot += c;
      colptr[m_begin + g + 1] = tot;
    }
    for (int64_t g = m_end + 1; g <= m; ++g) colptr[g] = tot;
    if (tot > capacity)
      rc = cdgpu_set_error(CDGPU_ECAP, "CSC output needs %lld entries, capacity is %lld", (long long)tot, (long long)capacity);
  }
  if (!rc && tot > 0) {
    e = cudaMallocAsync((void **)&drow, (size_t)tot * sizeof(long long), s);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&dval, (size_t)tot * sizeof(double), s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dcnt, cnt.data(), (size_t)mloc * sizeof(long long), cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) fail(e, "csc staging");
    if (!rc) {
      vc_csc_fill_kernel<<<blocks, 256, 0, s>>>(dout, (int)ep, (int)m_begin, (int)m_end, dcnt, drow, dval);
      CD_COUNT_LAUNCH(1);
      static_assert(sizeof(long long) == sizeof(int64_t), "rowval is copied as int64_t");
      e = cudaMemcpyAsync(rowval, drow, (size_t)tot * sizeof(int64_t), cudaMemcpyDeviceToHost, s);
      if (e == cudaSuccess) e = cudaMemcpyAsync(nzval, dval, (size_t)tot * sizeof(double), cudaMemcpyDeviceToHost, s);
      if (e == cudaSuccess) e = cudaStreamSynchronize(s);
      if (e != cudaSuccess) fail(e, "csc fill");
    }
  }
  for (void *ptr : {(void *)dcnt, (void *)drow, (void *)dval})
    if (ptr) cudaFreeAsync(ptr, s);
  return rc;
}

// host driver of the moment form; grid points are processed in chunks so the moment blocks stay <= ~1 GiB
const void *vc_cov_pick_std(int nu); // vc_cov_std.cu
const void *vc_cov_pick_lvo(int nu);  // vc_cov_lvo.cu

static int vc_solve_moment(const double *X, int64_t n, int64_t p, int64_t ldx, const double *z, const double *y,
                           const double *zgrid, int64_t m, int64_t m_begin, int64_t m_end, int degree, int kernel_kind,
                           double bandwidth, double lambda0, const cdgpu_options *opt, int device, double *out,
                           double *outR, cdgpu_stats *stats, const double *harr = nullptr, double *lvo_err = nullptr,
                           int64_t chain = 1, const VcCsc *csc = nullptr) {
  // csc != null: out / outR are null and the results leave the device as CSC (cdgpu_vc_solve_csc)
  const bool wantR = outR || (csc && csc->colptrR);
  // harr != null: the problems are those of lvocv_locpolyl1 (m = numH * n, zgrid unused, out/outR null, one squared
  // prediction error per problem into lvo_err)
  const int64_t ep = p * (degree + 1), mloc = m_end - m_begin, P2 = p * (p + 1) / 2, PA = P2 + p + 2;
  const int nq = 2 * degree + 1;
  const long long ldz = (n + 1) & ~(int64_t)1, ldc = (PA + 1) & ~(int64_t)1;
  int64_t chunk = (int64_t)((1ll << 30) / ((long long)ldc * nq * 8));
  chunk = std::max<int64_t>(1, std::min<int64_t>(chunk, mloc));
  chunk = std::min<int64_t>(chunk, std::max<int64_t>(1, (int64_t)((1ll << 30) / ((long long)ldz * nq * 8))));
  if (const char *env = getenv("CDGPU_VC_CHUNK")) chunk = std::max<int64_t>(1, std::min<int64_t>(chunk, atoll(env))); // tests: force several chunks
  // chains of warm-started grid points never straddle two chunks (one kernel launch each)
  chain = std::max<int64_t>(1, std::min<int64_t>(chain, mloc));
  if (chain > chunk)
    return cdgpu_set_error(CDGPU_ECAP, "chain of %lld grid points exceeds the %lld whose moment blocks fit one chunk", (long long)chain,
                           (long long)chunk);
  chunk = chunk / chain * chain;
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
  if (out || csc) VM_TRY(cudaMallocAsync((void **)&dout, (size_t)ep * m * sizeof(double), s));
  if (harr) {
    const int64_t numH = m / n;
    VM_TRY(cudaMallocAsync((void **)&dharr, (size_t)numH * sizeof(double), s));
    VM_TRY(cudaMallocAsync((void **)&derr, (size_t)m * sizeof(double), s));
    VM_TRY(cudaMemcpyAsync(dharr, harr, (size_t)numH * sizeof(double), cudaMemcpyHostToDevice, s));
  }
  VM_TRY(cudaMallocAsync((void **)&dst, (size_t)m * sizeof(DevStats), s));
  if (wantR) VM_TRY(cudaMallocAsync((void **)&doutR, (size_t)ep * m * sizeof(double), s));
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
  const int ring = 4; // columns of the compact Gram in flight per warp (see vc_cov_kernel.cuh)
  const void *kfn = harr ? vc_cov_pick_lvo(nu) : vc_cov_pick_std(nu);
  const size_t wsz = vc_cov_warp_bytes((int)ep, vc_cov_inst(nu), ring);
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
    a.chain = (int)chain;
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
  if (csc) {
    rc = vc_emit_csc(s, dout, ep, m, m_begin, m_end, csc->capacity, csc->colptr, csc->rowval, csc->nzval);
    if (!rc && csc->colptrR) rc = vc_emit_csc(s, doutR, ep, m, m_begin, m_end, csc->capacity, csc->colptrR, csc->rowvalR, csc->nzvalR);
    if (rc) {
      cleanup();
      return rc;
    }
  }
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
                         cdgpu_stats *stats, int64_t chain = 1, const VcCsc *csc = nullptr);
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
  if (ep > VC_COV_MAX_EP)
    return cdgpu_set_error(CDGPU_ECAP, "lvocv needs the moment form: p*(degree+1) = %lld exceeds %d", (long long)ep, VC_COV_MAX_EP);
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
API int cdgpu_vc_solve_chain(const double *X, int64_t n, int64_t p, int64_t ldx, const double *z, const double *y,
                             const double *zgrid, int64_t m, int64_t m_begin, int64_t m_end, int degree, int kernel_kind,
                             double bandwidth, double lambda0, const cdgpu_options *opt, int64_t chain, int device,
                             double *out, double *outR, cdgpu_stats *stats) {
  return api_guard([&]() -> int {
  return vc_solve_impl(X, n, p, ldx, z, y, zgrid, m, m_begin, m_end, degree, kernel_kind, bandwidth, lambda0, opt, device, out,
                       outR, stats, chain);
  });
}
API int cdgpu_vc_solve_csc(const double *X, int64_t n, int64_t p, int64_t ldx, const double *z, const double *y,
                           const double *zgrid, int64_t m, int64_t m_begin, int64_t m_end, int degree, int kernel_kind,
                           double bandwidth, double lambda0, const cdgpu_options *opt, int64_t chain, int device,
                           int64_t capacity, int64_t *colptr, int64_t *rowval, double *nzval, int64_t *colptrR,
                           int64_t *rowvalR, double *nzvalR, cdgpu_stats *stats) {
  return api_guard([&]() -> int {
  if (!colptr || capacity < 0 || (capacity > 0 && (!rowval || !nzval)) || (colptrR && capacity > 0 && (!rowvalR || !nzvalR)))
    return cdgpu_set_error(CDGPU_EARG, "null pointer");
  const VcCsc csc = {capacity, colptr, rowval, nzval, colptrR, rowvalR, nzvalR};
  return vc_solve_impl(X, n, p, ldx, z, y, zgrid, m, m_begin, m_end, degree, kernel_kind, bandwidth, lambda0, opt, device, nullptr,
                       nullptr, stats, chain, &csc);
  });
}
static int vc_solve_impl(const double *X, int64_t n, int64_t p, int64_t ldx, const double *z, const double *y,
                         const double *zgrid, int64_t m, int64_t m_begin, int64_t m_end, int degree, int kernel_kind,
                         double bandwidth, double lambda0, const cdgpu_options *opt, int device, double *out, double *outR,
                         cdgpu_stats *stats, int64_t chain, const VcCsc *csc) {
  if (!X || !z || !y || !zgrid || !opt || (!out && !csc)) return cdgpu_set_error(CDGPU_EARG, "null pointer");
  const bool wantR = outR || (csc && csc->colptrR);
  if (chain < 1) return cdgpu_set_error(CDGPU_EARG, "chain must be at least 1");
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
  if (mloc == 0) {
    if (csc) {
      for (int64_t g = 0; g <= m; ++g) csc->colptr[g] = 0;
      if (csc->colptrR)
        for (int64_t g = 0; g <= m; ++g) csc->colptrR[g] = 0;
    }
    return CDGPU_OK;
  }
  {
    // moment (covariance) form unless the expanded problem is too wide for the warp kernel's registers;
    // CDGPU_VC_FORM=naive keeps the residual-form kernels below
    const char *form = getenv("CDGPU_VC_FORM");
    // (256 < ep <= 512: the 12- and 16-slot instances of the moment kernel spill registers, so the plain solve keeps the
    // residual form there and only what needs the moment form - refit, chained grid points - takes them)
    if ((ep <= 256 && (wantR || !(form && strcmp(form, "naive") == 0))) || (ep <= VC_COV_MAX_EP && (wantR || chain > 1)))
      return vc_solve_moment(X, n, p, ldx, z, y, zgrid, m, m_begin, m_end, degree, kernel_kind, bandwidth, lambda0, opt,
                             device, out, outR, stats, nullptr, nullptr, chain, csc);
    if (chain > 1)
      return cdgpu_set_error(CDGPU_ECAP, "chained grid points need the moment form: p*(degree+1) = %lld exceeds %d",
                             (long long)ep, VC_COV_MAX_EP);
    if (wantR)
      return cdgpu_set_error(CDGPU_ECAP, "refit on the device needs the moment form: p*(degree+1) = %lld exceeds %d",
                             (long long)ep, VC_COV_MAX_EP);
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
  if (out)
    VC_TRY(cudaMemcpyAsync(out + m_begin * ep, dout + m_begin * ep, (size_t)mloc * ep * sizeof(double),
                           cudaMemcpyDeviceToHost, s));
  VC_TRY(cudaStreamSynchronize(s));
  if (csc) {
    rc = vc_emit_csc(s, dout, ep, m, m_begin, m_end, csc->capacity, csc->colptr, csc->rowval, csc->nzval);
    if (rc) {
      cleanup();
      return rc;
    }
  }
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
