// fp64_probe.cu — what the FP64 vector pipe of this GPU does: dependent-op latency (DADD, DMUL, DFMA,
// division, shuffle of a double) and per-SM throughput of independent DFMAs at 4..32 warps.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_probe fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void lat_kernel(double *out, long long *cyc, double a, double b, int iters) {
  double x = a, y = b;
  long long t0, t1;
  // DADD chain
  t0 = clock64();
  for (int i = 0; i < iters; ++i) x = __dadd_rn(x, y);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) x = __dmul_rn(x, y);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[1] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) x = fma(x, y, a);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[2] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) x = y / (x + 1.5);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[3] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) x = __shfl_sync(0xffffffffu, x, (i + 1) & 31) + y;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[4] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) x = (x > y ? x - y : (x < -y ? x + y : 0.0)) + a; // shrink + add
  t1 = clock64();
  if (threadIdx.x == 0) cyc[5] = t1 - t0;
  float f = (float)a, g = (float)b;
  t0 = clock64();
  for (int i = 0; i < iters; ++i) f = fmaf(f, g, 1.0f);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[6] = t1 - t0;
  out[threadIdx.x] = x + f;
}

__global__ void thr_kernel(double *out, long long *cyc, double a, double b, int iters) {
  double x0 = a, x1 = a + 1, x2 = a + 2, x3 = a + 3, x4 = a + 4, x5 = a + 5, x6 = a + 6, x7 = a + 7;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, b, a); x1 = fma(x1, b, a); x2 = fma(x2, b, a); x3 = fma(x3, b, a);
    x4 = fma(x4, b, a); x5 = fma(x5, b, a); x6 = fma(x6, b, a); x7 = fma(x7, b, a);
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

int main() {
  double *out;
  long long *cyc, h[8];
  cudaMalloc(&out, 148 * 1024 * 8);
  cudaMalloc(&cyc, 64);
  const int iters = 4096;
  lat_kernel<<<1, 32>>>(out, cyc, 1.0000001, 0.9999999, iters);
  cudaMemcpy(h, cyc, 56, cudaMemcpyDeviceToHost);
  const char *names[] = {"DADD", "DMUL", "DFMA", "DADD+DIV", "SHFL64+DADD", "shrink+DADD", "FFMA"};
  for (int i = 0; i < 7; ++i) printf("latency %-12s %.1f cycles/op\n", names[i], (double)h[i] / iters);
  for (int warps : {1, 2, 4, 8, 16, 32}) {
    thr_kernel<<<148, warps * 32>>>(out, cyc, 1.0000001, 0.9999999, iters);
    cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    double per_sm = (double)warps * 32 * 8 * iters / (double)h[0];
    printf("throughput %2d warps/SM: %.2f DFMA lanes/clk/SM  (%.2f TFLOP/s chip at 1.965 GHz)\n", warps, per_sm,
           per_sm * 2 * 148 * 1.965e9 / 1e12);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return 0;
}
