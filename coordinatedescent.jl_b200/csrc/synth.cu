// synth.cu — counter-based synthetic data for the benchmarks and the multi-GPU tests: element (i, j) of
// the design is a pure function of (seed, global row i, column j), so any row / column sharding over any number of
// GPUs (and a CPU sub-sample, oracle side: tests/helpers.py keyed_normal) sees exactly the same matrix.
// Box-Muller on two 32-bit halves of splitmix64(seed, i, j); evaluated in single precision, stored as double: the
// values only have to be reproducible, not exactly Gaussian.
#include "common.cuh"

namespace {

__global__ void synth_normal_kernel(double *out, long long rows, long long cols, long long ld, long long row0,
                                    long long col0, unsigned long long seed) {
  const long long total = rows * cols;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long j = e / rows, i = e - j * rows;
    const unsigned long long key =
        cd_splitmix(seed ^ cd_splitmix((unsigned long long)(row0 + i) * 0x9E3779B97F4A7C15ull + (unsigned long long)(col0 + j)));
    const unsigned u1 = (unsigned)(key >> 32), u2 = (unsigned)key;
    const float a = ((float)(u1 >> 8) + 0.5f) * (1.0f / 16777216.0f); // (0, 1)
    const float b = ((float)(u2 >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float r = sqrtf(-2.0f * logf(a));
    out[i + j * ld] = (double)(r * cospif(2.0f * b));
  }
}

} // namespace

extern "C" __attribute__((visibility("default"))) int cdgpu_synth_normal(double *d_out, int64_t rows, int64_t cols, int64_t ld,
                                                                          int64_t row0, int64_t col0, uint64_t seed, int device) {
  if (!d_out || rows < 0 || cols < 0 || ld < rows) return cdgpu_set_error(CDGPU_EARG, "bad arguments");
  CD_TRY(cd_use_device(device));
  if (rows == 0 || cols == 0) return CDGPU_OK;
  synth_normal_kernel<<<148 * 8, 256>>>(d_out, rows, cols, ld, row0, col0, seed);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaDeviceSynchronize());
  CD_COUNT_LAUNCH(1);
  return CDGPU_OK;
}
