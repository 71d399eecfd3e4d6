// refit.cu — post-selection least squares on a support: refitLassoPath (src/lasso.jl:208-225), `X[:, S] \ Y`.
// The normal equations on the support come from the handle's own device data: X_S'[W]X_S, X_S'[W]y by one warp per
// pair of columns (naive-form handles), or A[S,S], -b[S] gathered from the covariance-form handle; one CTA then
// solves the ns x ns system by a left-looking Cholesky in global scratch (columns contiguous: coalesced) and two
// triangular sweeps.  ns <= CD_GCAP = 4096 (the scratch of the handle's active Gram, common.cuh: cd_gram_cap).
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int RF_T = 512;

// M (ns x ns, ld) lower triangle + rhs from a naive-form handle: one warp per (i >= j) pair / rhs entry
__global__ void refit_gram_naive_kernel(const double *__restrict__ X, long long ldx, int n, const double *__restrict__ y,
                                        const double *__restrict__ w, const int *__restrict__ S, int ns, double *M, int ld,
                                        double *rhs) {
  const int lane = threadIdx.x & 31;
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long npair = (long long)ns * (ns + 1) / 2;
  for (long long idx = gw; idx < npair + ns; idx += nw) {
    const double *ci, *cj;
    int i = 0, j = 0;
    if (idx >= npair) {
      i = (int)(idx - npair);
      ci = X + (long long)S[i] * ldx;
      cj = y;
    } else {
      i = (int)((sqrt(8.0 * (double)idx + 1.0) - 1.0) * 0.5);
      while ((long long)i * (i + 1) / 2 > idx) --i;
      while ((long long)(i + 1) * (i + 2) / 2 <= idx) ++i;
      j = (int)(idx - (long long)i * (i + 1) / 2);
      ci = X + (long long)S[i] * ldx;
      cj = X + (long long)S[j] * ldx;
    }
    double s0 = 0.0, s1 = 0.0;
    int t = lane;
    for (; t + 32 < n; t += 64) {
      s0 = fma(w ? __ldg(ci + t) * __ldg(w + t) : __ldg(ci + t), __ldg(cj + t), s0);
      s1 = fma(w ? __ldg(ci + t + 32) * __ldg(w + t + 32) : __ldg(ci + t + 32), __ldg(cj + t + 32), s1);
    }
    for (; t < n; t += 32) s0 = fma(w ? __ldg(ci + t) * __ldg(w + t) : __ldg(ci + t), __ldg(cj + t), s0);
    const double v = warp_sum(s0 + s1);
    if (lane == 0) {
      if (idx >= npair)
        rhs[i] = v;
      else
        M[i + (long long)j * ld] = v;
    }
  }
}

__global__ void refit_gram_quad_kernel(const double *__restrict__ A, long long lda, const double *__restrict__ b,
                                       const int *__restrict__ S, int ns, double *M, int ld, double *rhs) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < (long long)ns * ns; idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % ns), j = (int)(idx / ns);
    if (i >= j) M[i + (long long)j * ld] = A[S[i] + (long long)S[j] * lda];
    if (j == 0) rhs[i] = -b[S[i]];
  }
}

// one CTA: Cholesky (lower, in place) + L y = rhs + L' x = y; flag[0] = 1 when not positive definite
__global__ void __launch_bounds__(RF_T, 1) refit_solve_kernel(double *M, int ld, int ns, double *rhs, int *flag) {
  __shared__ double lrow[CD_GCAP]; // row j of L while column j is formed (32 KB)
  __shared__ double red[RF_T / 32];
  __shared__ double sdiag;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  bool ok = true;
  for (int j = 0; j < ns; ++j) {
    for (int k = tid; k < j; k += RF_T) lrow[k] = M[j + (long long)k * ld];
    __syncthreads();
    // column j: rows i = j + tid, j + tid + T, ...
    for (int i = j + tid; i < ns; i += RF_T) {
      double acc = M[i + (long long)j * ld];
      int k = 0;
      for (; k + 4 <= j; k += 4) {
        const double a0 = __ldcg(M + i + (long long)k * ld), a1 = __ldcg(M + i + (long long)(k + 1) * ld);
        const double a2 = __ldcg(M + i + (long long)(k + 2) * ld), a3 = __ldcg(M + i + (long long)(k + 3) * ld);
        acc = fma(-a0, lrow[k], acc);
        acc = fma(-a1, lrow[k + 1], acc);
        acc = fma(-a2, lrow[k + 2], acc);
        acc = fma(-a3, lrow[k + 3], acc);
      }
      for (; k < j; ++k) acc = fma(-__ldcg(M + i + (long long)k * ld), lrow[k], acc);
      if (i == j) sdiag = acc;
      M[i + (long long)j * ld] = acc; // unscaled; scaled below once the pivot is known
    }
    __syncthreads();
    const double djj = sdiag;
    if (!(djj > 0.0)) ok = false;
    const double d = sqrt(djj);
    for (int i = j + tid; i < ns; i += RF_T) M[i + (long long)j * ld] = i == j ? d : M[i + (long long)j * ld] / d;
    __syncthreads();
  }
  for (int j = 0; j < ns; ++j) { // L y = rhs
    const double yj = rhs[j] / M[j + (long long)j * ld];
    __syncthreads();
    if (tid == 0) rhs[j] = yj;
    for (int i = j + 1 + tid; i < ns; i += RF_T) rhs[i] = fma(-M[i + (long long)j * ld], yj, rhs[i]);
    __syncthreads();
  }
  for (int j = ns - 1; j >= 0; --j) { // L' x = y
    double s = 0.0;
    for (int i = j + 1 + tid; i < ns; i += RF_T) s = fma(M[i + (long long)j * ld], rhs[i], s);
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int q = 0; q < RF_T / 32; ++q) t += red[q];
      rhs[j] = (rhs[j] - t) / M[j + (long long)j * ld];
    }
    __syncthreads();
  }
  if (tid == 0) flag[0] = ok ? 0 : 1;
}

} // namespace

// M: (ns x ns, ld = ns rounded up to even) + rhs behind it in `scratch`; dS: the support (0-based) on the device
int launch_refit(cdgpu_handle_s *h, const int *dS, int ns, double *scratch, int *flag) {
  const int ld = (ns + 1) & ~1;
  double *M = scratch, *rhs = scratch + (size_t)ld * ns;
  if (h->kind == CDGPU_LOSS_QUAD) {
    const long long tot = (long long)ns * ns;
    refit_gram_quad_kernel<<<(unsigned)std::min<long long>((tot + 255) / 256, 4096), 256, 0, h->stream>>>(h->dX, h->ld, h->dy, dS, ns, M,
                                                                                                         ld, rhs);
  } else {
    refit_gram_naive_kernel<<<h->sm_count * 4, 256, 0, h->stream>>>(h->dX, h->ld, (int)h->n, h->dy,
                                                                    h->kind == CDGPU_LOSS_WLS ? h->dw : nullptr, dS, ns, M, ld, rhs);
  }
  CUDA_TRY(cudaGetLastError());
  refit_solve_kernel<<<1, RF_T, 0, h->stream>>>(M, ld, ns, rhs, flag);
  CUDA_TRY(cudaGetLastError());
  CD_COUNT_LAUNCH(2);
  return CDGPU_OK;
}
