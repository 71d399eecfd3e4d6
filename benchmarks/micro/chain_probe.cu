// chain_probe.cu — the active-set chain engine alone (one CTA, synthetic compact G), with parts
// switched off to see what bounds it.  build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -DCHAIN_PROBE -I../../coordinatedescent.jl_b200/csrc -o chain_probe chain_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "chain_engine.cuh"

int cdgpu_set_error(int code, const char *, ...) { return code; }
long long g_cdgpu_launches = 0;

struct Pol { // covariance-form update with unit consts
  static constexpr bool HAS_RR = false;
  const double *b, *ainv;
  double lam;
  __device__ __forceinline__ void load_consts(int k, double &c0, double &c1, double &c2) const {
    c0 = __ldg(b + k);
    c1 = __ldg(ainv + k);
    c2 = __dmul_rn(c1, lam);
  }
  static constexpr bool FAST_V = true;
  __device__ __forceinline__ double enter(double g, double be, double c0, double c1) const { return __dsub_rn(be, __dmul_rn(g + c0, c1)); }
  __device__ __forceinline__ void step(double g, double be, double c0, double c1, double c2, double, double &nw, double &h,
                                       double &dr) const {
    const double v = __dsub_rn(be, __dmul_rn(g + c0, c1));
    nw = cd_shrink(v, c2);
    h = nw - be;
    dr = 0.0;
  }
  static __device__ __forceinline__ double apply(double g, double Gv, double h) { return __dadd_rn(g, __dmul_rn(Gv, h)); }
};

constexpr int T = 512, CAP = 2048;
__global__ void __launch_bounds__(T, 1) probe(const double *G, int m, const double *b, const double *ainv, double lam, int passes,
                                              long long *prof, double *beta_out, unsigned char *inlist, int spin_flag_unused) {
  extern __shared__ __align__(16) unsigned char raw[];
  if (blockIdx.x != 0) return;
  double *stage = reinterpret_cast<double *>(raw);
  double *g = stage + chain::STAGE_DOUBLES, *be = g + CAP;
  int *row = reinterpret_cast<int *>(be + CAP);
  unsigned short *ord = reinterpret_cast<unsigned short *>(row + CAP), *pos = ord + CAP;
  __shared__ chain::Shared sh;
  for (int i = threadIdx.x; i < m; i += T) {
    row[i] = i;
    be[i] = 0.0;
    g[i] = 0.0;
  }
  if (threadIdx.x < 8) prof[threadIdx.x] = 0;
  __syncthreads();
  chain::State S{m, row, row, g, be, ord, pos, stage, &sh, G, m, prof};
  Pol P{b, ainv, lam};
  long long t0 = clock64();
  chain::Result r = chain::run<T>(S, P, 0.0, passes, 0, true, 1, -1.0, inlist);
  long long t1 = clock64();
  for (int i = threadIdx.x; i < m; i += T) beta_out[i] = be[i];
  if (threadIdx.x == 0) {
    prof[8] = t1 - t0;
    prof[9] = r.accepted;
    prof[10] = r.visits;
  }
}

int main(int argc, char **argv) {
  const int m = argc > 1 ? atoi(argv[1]) : 828, passes = argc > 2 ? atoi(argv[2]) : 50;
  std::vector<double> G((size_t)m * m), b(m), ainv(m);
  srand(1);
  // G = I + small symmetric noise (diagonally dominant), b random
  for (int i = 0; i < m; ++i)
    for (int j = 0; j <= i; ++j) {
      double v = i == j ? 1.0 : 0.3 * (rand() / (double)RAND_MAX - 0.5) / 30.0;
      G[i + (size_t)j * m] = G[j + (size_t)i * m] = v;
    }
  for (int i = 0; i < m; ++i) {
    b[i] = rand() / (double)RAND_MAX - 0.5;
    ainv[i] = 1.0;
  }
  double *dG, *db, *da, *dbeta;
  long long *dprof;
  unsigned char *dinl;
  cudaMalloc(&dG, G.size() * 8);
  cudaMalloc(&db, m * 8);
  cudaMalloc(&da, m * 8);
  cudaMalloc(&dbeta, m * 8);
  cudaMalloc(&dprof, 16 * 8);
  cudaMalloc(&dinl, m);
  cudaMemcpy(dG, G.data(), G.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), m * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(da, ainv.data(), m * 8, cudaMemcpyHostToDevice);
  const size_t dyn = chain::STAGE_DOUBLES * 8 + CAP * (8 + 8 + 4 + 2 + 2);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
  for (int mode = 0; mode < 8; ++mode) {
    cudaMemcpyToSymbol(chain::probe_mode, &mode, sizeof(int));
    probe<<<1, T, dyn>>>(dG, m, db, da, 0.01, passes, dprof, dbeta, dinl, 0);
    long long pf[16];
    cudaMemcpy(pf, dprof, sizeof pf, cudaMemcpyDeviceToHost);
    const double nblk = (double)passes * ((m + 31) / 32);
    printf("mode %d (skip: %s%s%s) total %.0f cyc/block | warp0 panel %.0f chain %.0f barrier %.0f passend %.0f | workers stage %.0f apply %.0f "
           "barrier %.0f | accepted %lld visits %lld\n",
           mode, mode & 1 ? "apply " : "", mode & 2 ? "stage " : "", mode & 4 ? "chain" : "", pf[8] / nblk, pf[0] / nblk, pf[1] / nblk,
           pf[2] / nblk, pf[3] / nblk, pf[4] / nblk, pf[5] / nblk, pf[6] / nblk, pf[9], pf[10]);
  }
  printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
