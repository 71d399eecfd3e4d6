// lazy_gram.cu — the covariance form WITHOUT forming all of A = X'X/n.
//
// The covariance-form step (src/cd_differentiable_function.jl:324-348) only ever reads diag(A), b = -X'y/n and the
// columns A[:,k] of coordinates that are (or become) non-zero — at BASELINE C2 about 10^2 of 2*10^4 columns over a
// 100-lambda path.  A lazy handle therefore forms, up front, only diag(A) and b (one pass over X, HBM bound) and
// the columns on demand in batches of 128: A[:, S] = X' X[:, S] / n is a skinny FP64 tensor-core GEMM through the
// same DMMA kernel as the full Gram (gram_dmma.cu: launch_gemm_tn), written into slots of a column cache.  The sweep
// kernel (cov_sweep.cu) addresses column k through the slot map and leaves at a consistent point when a coordinate
// enters whose column is missing; the host loop in api.cu forms a batch (that column plus the currently
// highest-scoring |grad_j|/omega_j candidates, the ones most likely to enter next) and resumes the kernel.
// glmnet's "covariance updates" are the same idea on a CPU.
#include <algorithm>

#include "common.cuh"

namespace {

// diag_j = sum_i X_ij^2 / n, b_j = -(X_j'y) / n, ainv_j = 1/diag_j: one warp per column, the column is read once
__global__ void diag_xty_kernel(const double *__restrict__ X, long long n, int p, long long ldx,
                                const double *__restrict__ y, const double *__restrict__ w, double divisor, double *diag,
                                double *b, double *ainv, int accumulate, int finish) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int k = blockIdx.x * wpb + (threadIdx.x >> 5); k < p; k += gridDim.x * wpb) {
    const double *col = X + (long long)k * ldx;
    double s0 = 0.0, s1 = 0.0, t0 = 0.0, t1 = 0.0;
    long long i = lane;
    if (w) {
      for (; i < n; i += 32) {
        const double x0 = __ldg(col + i), wx = __ldg(w + i) * x0;
        s0 = fma(wx, x0, s0);
        t0 = fma(wx, __ldg(y + i), t0);
      }
    }
    for (; i + 32 < n; i += 64) {
      const double x0 = __ldg(col + i), x1 = __ldg(col + i + 32);
      s0 = fma(x0, x0, s0);
      s1 = fma(x1, x1, s1);
      t0 = fma(x0, __ldg(y + i), t0);
      t1 = fma(x1, __ldg(y + i + 32), t1);
    }
    for (; i < n; i += 32) {
      const double x0 = __ldg(col + i);
      s0 = fma(x0, x0, s0);
      t0 = fma(x0, __ldg(y + i), t0);
    }
    double s = warp_sum(s0 + s1), t = warp_sum(t0 + t1);
    if (lane == 0) {
      if (accumulate) { // row-chunked input (host staging): raw sums accumulate, the last chunk finishes
        s += diag[k];
        t = b[k] - t;
      } else {
        t = -t;
      }
      if (finish) {
        s = s / divisor;
        t = t / divisor;
        ainv[k] = 1.0 / s;
      }
      diag[k] = s;
      b[k] = t;
    }
  }
}

// score_j = |Ax_j + b_j| / omega_j for columns not formed yet, -1 for formed ones: who is likely to enter next
__global__ void lazy_score_kernel(const double *Ax, const double *b, const double *omega, const int *slot, int p, double *out) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < p; j += gridDim.x * blockDim.x) {
    double s = fabs(Ax[j] + b[j]);
    if (omega) s = s / omega[j];
    out[j] = slot[j] >= 0 ? -1.0 : (s == s ? s : 0.0);
  }
}

// B[:, q] = X[:, idx[q]] (q < nb; columns beyond nb up to nbpad are zero-filled), and slot[idx[q]] = slot0 + q
__global__ void gather_cols_kernel(const double *__restrict__ X, long long ldx, long long n, const double *__restrict__ w,
                                   const int *__restrict__ idx, int nb, int nbpad, double *B, long long ldb, int *slot,
                                   int slot0) {
  const int q = blockIdx.y;
  const bool live = q < nb;
  const double *src = X + (long long)(live ? idx[q] : 0) * ldx;
  double *dst = B + (long long)q * ldb;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < ldb; i += (long long)gridDim.x * blockDim.x)
    dst[i] = (live && i < n) ? (w ? w[i] * src[i] : src[i]) : 0.0;
  if (slot && live && blockIdx.x == 0 && threadIdx.x == 0) slot[idx[q]] = slot0 + q;
  (void)nbpad;
}

__global__ void assign_slots_kernel(const int *idx, int nb, int slot0, int *slot) {
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nb; q += gridDim.x * blockDim.x) slot[idx[q]] = slot0 + q;
}
__global__ void fill_int_kernel(int *a, int n, int v) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) a[i] = v;
}
__global__ void sqrt_vec_kernel(const double *a, int n, double *out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = sqrt(a[i]);
}

} // namespace

int launch_diag_xty(cdgpu_handle_s *h, const double *X, long long n, int p, long long ldx, const double *y, const double *w,
                    double divisor, double *diag, double *b, double *ainv, int accumulate, int finish) {
  diag_xty_kernel<<<min((p + 7) / 8, h->sm_count * 8), 256, 0, h->stream>>>(X, n, p, ldx, y, w, divisor, diag, b, ainv,
                                                                            accumulate, finish);
  CUDA_TRY(cudaGetLastError());
  CD_COUNT_LAUNCH(1);
  return CDGPU_OK;
}
int launch_lazy_score(cdgpu_handle_s *h, const double *Ax, const double *b, const double *omega, const int *slot, int p,
                      double *out) {
  lazy_score_kernel<<<(p + 255) / 256, 256, 0, h->stream>>>(Ax, b, omega, slot, p, out);
  CUDA_TRY(cudaGetLastError());
  CD_COUNT_LAUNCH(1);
  return CDGPU_OK;
}
int launch_assign_slots(cudaStream_t stream, const int *idx, int nb, int slot0, int *slot) {
  assign_slots_kernel<<<1, 128, 0, stream>>>(idx, nb, slot0, slot);
  CUDA_TRY(cudaGetLastError());
  CD_COUNT_LAUNCH(1);
  return CDGPU_OK;
}
int launch_gather_cols(cudaStream_t stream, const double *X, long long ldx, long long n, const double *w, const int *idx, int nb,
                       int nbpad, double *B, long long ldb, int *slot, int slot0) {
  dim3 grid((unsigned)std::min<long long>((ldb + 255) / 256, 64), (unsigned)nbpad);
  gather_cols_kernel<<<grid, 256, 0, stream>>>(X, ldx, n, w, idx, nb, nbpad, B, ldb, slot, slot0);
  CUDA_TRY(cudaGetLastError());
  CD_COUNT_LAUNCH(1);
  return CDGPU_OK;
}
int launch_fill_int(cdgpu_handle_s *h, int *a, int n, int v) {
  fill_int_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(a, n, v);
  CUDA_TRY(cudaGetLastError());
  CD_COUNT_LAUNCH(1);
  return CDGPU_OK;
}
int launch_sqrt_vec(cdgpu_handle_s *h, const double *a, int n, double *out) {
  sqrt_vec_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(a, n, out);
  CUDA_TRY(cudaGetLastError());
  CD_COUNT_LAUNCH(1);
  return CDGPU_OK;
}
