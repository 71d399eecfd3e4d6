// gram_dmma.cu — K1: G = X'X / n (exactly symmetric) and c = -X'y / n for the covariance form
// (what the reference's users write by hand: test/lasso.jl:48,88), as an FP64 tensor-core SYRK.
//
// sm_100a has no tcgen05 kind for f64 (ptxas: "Unknown modifier .kind::f64"), so the FP64 tensor
// path is the warp-level mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4).  Design:
//   * X is column-major n x p, so both operands of X'X are K-contiguous: a tile is 128 columns of X
//     by BK = 16 rows, i.e. 128 runs of 128 contiguous bytes -> 16-byte cp.async (LDGSTS), 4 stages.
//   * CTA tile 128 x 128, 8 warps as 2 x 4, warp tile 64 x 32 = 8 x 4 DMMA fragments (64 accumulator
//     doubles / thread).  Shared rows are padded to 20 doubles so fragment loads are conflict free.
//   * operand fragments are double-buffered in registers ACROSS the k-tile barrier (the wait covers two
//     stages), so the DMMA pipe does not drain at each __syncthreads: 128.6 ms -> 113.4 ms at C2
//     (85 % -> ~99 % of the cuBLAS Dgemm rate).  CDGPU_GRAM_VARIANT selects the other tilings tried.
//   * only tiles on or below the diagonal are computed; the epilogue writes acc/n to (i,j) and the
//     same value to (j,i), so issymmetric(G) (cd_differentiable_function.jl:306) holds bit for bit.
//   * persistent CTAs walk a host-built tile list ordered in 12 x 12 super-tiles so the ~148 tiles in
//     flight share ~24 column panels of X through L2 (HBM traffic ~ 20 GB instead of ~130 GB at C2).
//   * the row-sharded multi-GPU form runs the same kernel on n_local rows without the 1/n, sums the
//     partial G|c buffer with one ncclAllReduce (nccl_comm.cu) and then scales.
#include <vector>

#include "common.cuh"

namespace {

constexpr int BM = 128, BN_FULL = 128;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N));
}
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// load one stage: rows (columns of X) [col0, col0+128) x k in [k0, k0+BK) into dst[128][LDK]
template <bool ALIGNED16, int BK, int LDK, int GT, int ROWS = 128>
__device__ __forceinline__ void load_tile(double *dst, const double *X, long long ldx, long long n, int p, int col0,
                                          long long k0, int tid) {
  if (ALIGNED16) {
    constexpr int CPR = BK / 2; // 16-byte chunks per row
    constexpr int ITERS = (ROWS * CPR + GT - 1) / GT;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int idx = tid + it * GT;
      if ((ROWS * CPR) % GT != 0 && idx >= ROWS * CPR) break;
      const int row = idx / CPR, ch = idx % CPR;
      const int col = col0 + row;
      const long long k = k0 + ch * 2;
      int bytes = 0;
      if (col < p && k < n) bytes = (k + 1 < n) ? 16 : 8;
      const double *src = X + (bytes ? ((long long)col * ldx + k) : 0);
      cp_async16(dst + row * LDK + ch * 2, src, bytes);
    }
  } else {
    constexpr int ITERS = (ROWS * BK + GT - 1) / GT;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int idx = tid + it * GT;
      if ((ROWS * BK) % GT != 0 && idx >= ROWS * BK) break;
      const int row = idx / BK, ch = idx % BK;
      const int col = col0 + row;
      const long long k = k0 + ch;
      const int bytes = (col < p && k < n) ? 8 : 0;
      const double *src = X + (bytes ? ((long long)col * ldx + k) : 0);
      cp_async8(dst + row * LDK + ch, src, bytes);
    }
  }
}

// WM x WN warps; warp tile (128/WM) x (128/WN) = MI x NI fragments of 8 x 8
// GEMM = false: SYRK on X (lower tiles, mirrored).  GEMM = true: C = X'B for a second K-contiguous operand
// B (n x pb, ldb): every tile of the pa x pb rectangle, plain stores (used by the batched
// varying-coefficient path, vc_batch.cu: all local Gram matrices in one FP64 tensor-core GEMM).
template <bool ALIGNED16, int WM, int WN, int BK, int STAGES, bool PIPE, bool GEMM = false, int BN = 128>
__global__ void __launch_bounds__(WM *WN * 32, 1)
    gram_syrk_kernel(const double *X, long long n, int p, long long ldx, double *G,
                     long long ldg, const int2 *__restrict__ tiles, int ntiles, double divisor, int mode,
                     const double *__restrict__ Bm = nullptr, int pb = 0, long long ldb = 0, long long kchunk = 0,
                     double *__restrict__ slab = nullptr, int slab_tiles = 0) {
  const double *XB = GEMM ? Bm : X;
  const int pB = GEMM ? pb : p;
  const long long ldB = GEMM ? ldb : ldx;
  constexpr int GT = WM * WN * 32, LDK = BK + 4, STAGE_DOUBLES = (BM + BN) * LDK;
  constexpr int MI = BM / (WM * 8), NI = BN / (WN * 8);
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp / WN, wn = warp % WN;
  const int g = lane >> 2, q = lane & 3;
  const long long n_all = n;

  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    // work item: tile (bi >= bj) and, in the row-split form (kchunk > 0, tall-skinny X), the chunk of rows it covers;
    // its partial sums go to slab[chunk][tile index] and are reduced in a fixed order afterwards
    const int bi = tiles[t].x & 0xffff, chunk = tiles[t].x >> 16, bj = tiles[t].y & 0xffff, tidx = tiles[t].y >> 16;
    const long long kbeg = kchunk > 0 ? (long long)chunk * kchunk : 0;
    X += kbeg;   // rows [kbeg, kend) of both operands
    XB += kbeg;
    n = kchunk > 0 ? min(n_all - kbeg, kchunk) : n_all;
    const long long nK = (n + BK - 1) / BK;
    const int colA = bi * BM, colB = bj * BN;
    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
      for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // prologue
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
      if (s < nK) {
        double *st = smem + s * STAGE_DOUBLES;
        load_tile<ALIGNED16, BK, LDK, GT>(st, X, ldx, n, p, colA, (long long)s * BK, tid);
        load_tile<ALIGNED16, BK, LDK, GT, BN>(st + BM * LDK, XB, ldB, n, pB, colB, (long long)s * BK, tid);
      }
      cp_commit();
    }
    if (!PIPE) {
      for (long long kt = 0; kt < nK; ++kt) {
        cp_wait<STAGES - 2>();
        __syncthreads();
        { // prefetch stage kt + STAGES - 1 into the slot freed by iteration kt - 1
          const long long kn = kt + STAGES - 1;
          if (kn < nK) {
            double *st = smem + (kn % STAGES) * STAGE_DOUBLES;
            load_tile<ALIGNED16, BK, LDK, GT>(st, X, ldx, n, p, colA, kn * BK, tid);
            load_tile<ALIGNED16, BK, LDK, GT, BN>(st + BM * LDK, XB, ldB, n, pB, colB, kn * BK, tid);
          }
          cp_commit();
        }
        const double *As = smem + (kt % STAGES) * STAGE_DOUBLES + (wm * MI * 8) * LDK;
        const double *Bs = smem + (kt % STAGES) * STAGE_DOUBLES + BM * LDK + (wn * NI * 8) * LDK;
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
          double af[MI], bf[NI];
#pragma unroll
          for (int i = 0; i < MI; ++i) af[i] = As[(i * 8 + g) * LDK + kk + q];
#pragma unroll
          for (int j = 0; j < NI; ++j) bf[j] = Bs[(j * 8 + g) * LDK + kk + q];
#pragma unroll
          for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NI; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
      }
    } else {
      // Register double-buffered fragments carried ACROSS the k-tile barrier: the wait at the top of
      // iteration kt covers stages kt and kt+1, so the first fragments of tile kt+1 are loaded while the
      // last DMMAs of tile kt issue, and the tensor pipe does not drain at every __syncthreads.
      double af[2][MI], bf[2][NI];
      auto load_frag = [&](int buf, long long kt, int kk) {
        const double *As = smem + (kt % STAGES) * STAGE_DOUBLES + (wm * MI * 8) * LDK;
        const double *Bs = smem + (kt % STAGES) * STAGE_DOUBLES + BM * LDK + (wn * NI * 8) * LDK;
#pragma unroll
        for (int i = 0; i < MI; ++i) af[buf][i] = As[(i * 8 + g) * LDK + kk + q];
#pragma unroll
        for (int j = 0; j < NI; ++j) bf[buf][j] = Bs[(j * 8 + g) * LDK + kk + q];
      };
      cp_wait<STAGES - 2>();
      __syncthreads();
      load_frag(0, 0, 0);
      for (long long kt = 0; kt < nK; ++kt) {
        cp_wait<(STAGES >= 3 ? STAGES - 3 : 0)>(); // stages kt and kt+1 have landed (this thread's part)
        __syncthreads();                           // ... and everyone's; slot of tile kt-1 is free
        {
          const long long kn = kt + STAGES - 1;
          if (kn < nK) {
            double *st = smem + (kn % STAGES) * STAGE_DOUBLES;
            load_tile<ALIGNED16, BK, LDK, GT>(st, X, ldx, n, p, colA, kn * BK, tid);
            load_tile<ALIGNED16, BK, LDK, GT, BN>(st + BM * LDK, XB, ldB, n, pB, colB, kn * BK, tid);
          }
          cp_commit();
        }
#pragma unroll
        for (int ks = 0; ks < BK / 4; ++ks) {
          const int cur = ks & 1;
          if (ks + 1 < BK / 4)
            load_frag(cur ^ 1, kt, (ks + 1) * 4);
          else if (kt + 1 < nK)
            load_frag(cur ^ 1, kt + 1, 0);
#pragma unroll
          for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int j = 0; j < NI; ++j) dmma(acc[i][j][0], acc[i][j][1], af[cur][i], bf[cur][j]);
        }
      }
    }
    cp_wait<0>();
    __syncthreads(); // all warps done with the last stage before the next tile's prologue overwrites it

    // epilogue: G[colA + m, colB + nn] and its mirror.  mode 0: store raw sums, 1: store sums / divisor,
    // 2: add the raw sums to what is there (row-chunked accumulation; both triangles hold the same value),
    // 3: as 2, then divide by divisor (last row chunk)
    if (kchunk > 0) { // row-split: raw partial tile to the slab (the reduction applies scale / mirror / diagonal rule)
      double *st = slab + ((size_t)chunk * slab_tiles + tidx) * (size_t)(BM * BN);
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) {
          const int r = wm * MI * 8 + i * 8 + g, cc = wn * NI * 8 + j * 8 + 2 * q;
          *reinterpret_cast<double2 *>(st + (size_t)r * BN + cc) = make_double2(acc[i][j][0], acc[i][j][1]);
        }
      X -= kbeg;
      XB -= kbeg;
      continue;
    }
    auto put = [&](int r, int cidx, double v) {
      if (mode >= 2) v += G[r + (long long)cidx * ldg];
      if (mode == 3) v = v / divisor;
      G[r + (long long)cidx * ldg] = v;
      G[cidx + (long long)r * ldg] = v;
    };
#pragma unroll
    for (int i = 0; i < MI; ++i) {
      const int row = colA + wm * MI * 8 + i * 8 + g;
#pragma unroll
      for (int j = 0; j < NI; ++j) {
        const int col = colB + wn * NI * 8 + j * 8 + 2 * q;
        double v0 = acc[i][j][0], v1 = acc[i][j][1];
        if (mode == 1) {
          v0 = v0 / divisor;
          v1 = v1 / divisor;
        }
        if (GEMM) {
          if (row < p) {
            if (col < pB) G[row + (long long)col * ldg] = v0;
            if (col + 1 < pB) G[row + (long long)(col + 1) * ldg] = v1;
          }
        } else if (row < p) {
          // diagonal tiles: the lower part (and the diagonal) is authoritative
          if (col < p && (bi != bj || row >= col)) put(row, col, v0);
          if (col + 1 < p && (bi != bj || row >= col + 1)) put(row, col + 1, v1);
        }
      }
    }
  }
}

// row-split form: G tile = sum over row chunks of the slab partials (fixed order: deterministic), then the same
// store rule as the direct epilogue (mode 0 raw / 1 divided; lower part authoritative on diagonal tiles; mirrored)
__global__ void gram_reduce_slabs_kernel(const double *__restrict__ slab, int nchunks, int slab_tiles, const int2 *__restrict__ base_tiles,
                                         int p, double *G, long long ldg, double divisor, int mode) {
  const int tidx = blockIdx.x;
  const int bi = base_tiles[tidx].x, bj = base_tiles[tidx].y;
  for (int e = threadIdx.x; e < BM * BN_FULL; e += blockDim.x) {
    const int r = e / BN_FULL, cc = e % BN_FULL;
    const int row = bi * BM + r, col = bj * BN_FULL + cc;
    if (row >= p || col >= p || (bi == bj && row < col)) continue;
    double v = 0.0;
    for (int c = 0; c < nchunks; ++c) v += slab[((size_t)c * slab_tiles + tidx) * (size_t)(BM * BN_FULL) + e];
    if (mode == 1) v = v / divisor;
    G[row + (long long)col * ldg] = v;
    G[col + (long long)row * ldg] = v;
  }
}

// c_j = -(X_j'y) [/ n]; one warp per column
__global__ void xty_kernel(const double *__restrict__ X, long long n, int p, long long ldx,
                           const double *__restrict__ y, double *c, double divisor, int mode) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int k = blockIdx.x * wpb + (threadIdx.x >> 5); k < p; k += gridDim.x * wpb) {
    const double *col = X + (long long)k * ldx;
    double s0 = 0.0, s1 = 0.0;
    long long i = lane;
    for (; i + 32 < n; i += 64) {
      s0 = fma(__ldg(col + i), __ldg(y + i), s0);
      s1 = fma(__ldg(col + i + 32), __ldg(y + i + 32), s1);
    }
    for (; i < n; i += 32) s0 = fma(__ldg(col + i), __ldg(y + i), s0);
    double s = warp_sum(s0 + s1);
    if (lane == 0) c[k] = mode == 1 ? -s / divisor : (mode == 2 ? c[k] - s : (mode == 3 ? (c[k] - s) / divisor : -s));
  }
}

// row-split GEMM form: C tile = sum over row chunks of the slab partials (fixed order), / divisor
__global__ void gemm_reduce_slabs_kernel(const double *__restrict__ slab, int nchunks, int slab_tiles, const int2 *__restrict__ base_tiles,
                                         int pa, int pb, double *C, long long ldc, double divisor, int bn) {
  const int tidx = blockIdx.x;
  const int bi = base_tiles[tidx].x, bj = base_tiles[tidx].y;
  for (int e = threadIdx.x; e < BM * bn; e += blockDim.x) {
    const int r = e % BM, cc = e / BM; // consecutive threads -> consecutive rows of one output column
    const int row = bi * BM + r, col = bj * bn + cc;
    if (row >= pa || col >= pb) continue;
    double v = 0.0;
    for (int c = 0; c < nchunks; ++c) v += slab[((size_t)c * slab_tiles + tidx) * (size_t)(BM * bn) + (size_t)r * bn + cc];
    C[row + (long long)col * ldc] = v / divisor;
  }
}

__global__ void scale_kernel(double *G, long long count, double divisor) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
    G[i] = G[i] / divisor;
}
// after a multi-rank allreduce the two copies (i,j) / (j,i) of an entry may have been summed in different orders
// (they sit in different chunks of the collective): the lower triangle is authoritative, as in the SYRK epilogue, so
// issymmetric(A) (cd_differentiable_function.jl:306) holds bit for bit on every rank.  32 x 32 tiles through shared memory.
__global__ void mirror_lower_kernel(double *G, long long ldg, int p) {
  __shared__ double t[32][33];
  const int bi = blockIdx.x, bj = blockIdx.y; // tile (rows bi, cols bj), bi >= bj: lower part
  if (bj > bi) return;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int r = ty; r < 32; r += blockDim.y) {
    const int i = bi * 32 + tx, j = bj * 32 + r;
    t[r][tx] = (i < p && j < p) ? G[i + (long long)j * ldg] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += blockDim.y) {
    const int i = bj * 32 + tx, j = bi * 32 + r; // upper element (i, j) = lower element (j, i) = t[tx][r]
    if (i < p && j < p && i < j) G[i + (long long)j * ldg] = t[tx][r];
  }
}

} // namespace

int launch_gram(cdgpu_handle_s *h, const double *X, long long n, int p, long long ldx, const double *y, double *G,
                double *c, double divisor, int mode) {
  const int nb = (p + BM - 1) / BM;
  // tile list in 12 x 12 super-tiles over the lower triangle
  const int S = 12;
  std::vector<int2> tiles;
  if (!h->dtiles) {
    tiles.reserve((size_t)nb * (nb + 1) / 2);
    for (int SI = 0; SI < nb; SI += S)
      for (int SJ = 0; SJ <= SI; SJ += S)
        for (int bi = SI; bi < min(SI + S, nb); ++bi)
          for (int bj = SJ; bj < min(SJ + S, nb) && bj <= bi; ++bj) tiles.push_back(make_int2(bi, bj));
    h->ntiles = (int)tiles.size();
  }
  const int ntiles = h->ntiles;
  if (!h->dtiles) { // built once per handle (depends only on p)
    // NB: stream-ordered copy + sync.  A plain cudaMemcpy from pageable memory may return before the DMA has
    // landed, and the kernel below runs on a non-blocking stream that does not order against it.
    CUDA_TRY(cudaMallocAsync((void **)&h->dtiles, (size_t)ntiles * sizeof(int2), h->stream)); // pool: no device sync
    CUDA_TRY(cudaMemcpyAsync(h->dtiles, tiles.data(), (size_t)ntiles * sizeof(int2), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream)); // `tiles` is a host temporary
  }
  const int2 *dtiles = static_cast<const int2 *>(h->dtiles);
  const bool aligned = ((reinterpret_cast<uintptr_t>(X) & 15) == 0) && ((ldx & 1) == 0);
  const long long ldg = ((long long)p + 1) & ~1ll;
  // Tall-skinny X (few tiles, many rows): with one work item per tile the last wave is mostly idle and there are
  // fewer tiles than SMs times two; split the ROWS as well: items = tiles x row chunks with nchunks chosen so that
  // the item count is a multiple of the SM count, partial tiles to a slab, one deterministic reduction.
  if (aligned && (mode == 0 || mode == 1) && ntiles < 2 * h->sm_count && n >= 65536 && !getenv("CDGPU_GRAM_NO_ROWSPLIT")) {
    auto gcd = [](int a_, int b_) {
      while (b_) {
        const int t_ = a_ % b_;
        a_ = b_;
        b_ = t_;
      }
      return a_;
    };
    int nchunks = h->sm_count / gcd(ntiles, h->sm_count);
    while ((long long)nchunks * ntiles < 6ll * h->sm_count) nchunks *= 2; // at least ~6 waves
    long long kchunk = ((n + nchunks - 1) / nchunks + 15) & ~15ll;
    if (kchunk < 4096) kchunk = 4096;
    nchunks = (int)((n + kchunk - 1) / kchunk);
    std::vector<int2> base((size_t)ntiles), items;
    {
      std::vector<int2> all;
      for (int SI = 0; SI < nb; SI += S)
        for (int SJ = 0; SJ <= SI; SJ += S)
          for (int bi = SI; bi < min(SI + S, nb); ++bi)
            for (int bj = SJ; bj < min(SJ + S, nb) && bj <= bi; ++bj) all.push_back(make_int2(bi, bj));
      base = all;
    }
    items.reserve((size_t)ntiles * nchunks);
    for (int ch = 0; ch < nchunks; ++ch)
      for (int t = 0; t < ntiles; ++t) items.push_back(make_int2(base[t].x | (ch << 16), base[t].y | (t << 16)));
    int2 *ditems = nullptr, *dbase = nullptr;
    double *slab = nullptr;
    const size_t slab_doubles = (size_t)nchunks * ntiles * BM * BN_FULL;
    CUDA_TRY(cudaMallocAsync((void **)&ditems, items.size() * sizeof(int2), h->stream));
    CUDA_TRY(cudaMallocAsync((void **)&dbase, base.size() * sizeof(int2), h->stream));
    CUDA_TRY(cudaMallocAsync((void **)&slab, slab_doubles * sizeof(double), h->stream));
    CUDA_TRY(cudaMemcpyAsync(ditems, items.data(), items.size() * sizeof(int2), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(dbase, base.data(), base.size() * sizeof(int2), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream)); // host temporaries
    auto kern = gram_syrk_kernel<true, 2, 4, 16, 4, true>;
    const size_t dyn = (size_t)4 * (BM + BN_FULL) * (16 + 4) * sizeof(double);
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    kern<<<min((int)items.size(), h->sm_count), 256, dyn, h->stream>>>(X, n, p, ldx, G, ldg, ditems, (int)items.size(), divisor, mode,
                                                                      nullptr, 0, 0, kchunk, slab, ntiles);
    CUDA_TRY(cudaGetLastError());
    gram_reduce_slabs_kernel<<<ntiles, 512, 0, h->stream>>>(slab, nchunks, ntiles, dbase, p, G, ldg, divisor, mode);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaFreeAsync(slab, h->stream));
    CUDA_TRY(cudaFreeAsync(ditems, h->stream));
    CUDA_TRY(cudaFreeAsync(dbase, h->stream));
    xty_kernel<<<min((p + 7) / 8, h->sm_count * 8), 256, 0, h->stream>>>(X, n, p, ldx, y, c, divisor, mode);
    CUDA_TRY(cudaGetLastError());
    CD_COUNT_LAUNCH(3);
    return CDGPU_OK;
  }
  const int grid = min(ntiles, h->sm_count);
  int variant = 5; // 2x4 warps, BK 16, 4 stages, fragments double-buffered across the k-tile barrier
  if (const char *env = getenv("CDGPU_GRAM_VARIANT")) variant = atoi(env);
  auto launch = [&](auto kern, int threads, size_t dyn) -> int {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    kern<<<grid, threads, dyn, h->stream>>>(X, n, p, ldx, G, ldg, dtiles, ntiles, divisor, mode, nullptr, 0, 0, 0, nullptr, 0);
    CUDA_TRY(cudaGetLastError());
    return CDGPU_OK;
  };
#define GRAM_DYN(BK, ST) ((size_t)(ST) * (BM + BN_FULL) * ((BK) + 4) * sizeof(double))
  int rc;
  if (!aligned)
    rc = launch(gram_syrk_kernel<false, 2, 4, 16, 4, false>, 256, GRAM_DYN(16, 4));
  else if (variant == 1)
    rc = launch(gram_syrk_kernel<true, 4, 4, 16, 4, false>, 512, GRAM_DYN(16, 4));
  else if (variant == 2)
    rc = launch(gram_syrk_kernel<true, 4, 4, 32, 3, false>, 512, GRAM_DYN(32, 3));
  else if (variant == 3)
    rc = launch(gram_syrk_kernel<true, 2, 4, 32, 3, false>, 256, GRAM_DYN(32, 3));
  else if (variant == 5)
    rc = launch(gram_syrk_kernel<true, 2, 4, 16, 4, true>, 256, GRAM_DYN(16, 4));
  else if (variant == 6)
    rc = launch(gram_syrk_kernel<true, 2, 4, 32, 3, true>, 256, GRAM_DYN(32, 3));
  else if (variant == 7)
    rc = launch(gram_syrk_kernel<true, 2, 4, 16, 5, true>, 256, GRAM_DYN(16, 5));
  else if (variant == 4)
    rc = launch(gram_syrk_kernel<true, 4, 2, 16, 4, false>, 256, GRAM_DYN(16, 4));
  else if (variant == 0)
    rc = launch(gram_syrk_kernel<true, 2, 4, 16, 4, false>, 256, GRAM_DYN(16, 4));
  else
    rc = launch(gram_syrk_kernel<true, 2, 4, 16, 4, true>, 256, GRAM_DYN(16, 4));
#undef GRAM_DYN
  CD_TRY(rc);
  xty_kernel<<<min((p + 7) / 8, h->sm_count * 8), 256, 0, h->stream>>>(X, n, p, ldx, y, c, divisor, mode);
  CUDA_TRY(cudaGetLastError());
  CD_COUNT_LAUNCH(2);
  return CDGPU_OK;
}

// C (pa x pb, ldc) = A'B / divisor for K-contiguous A (n x pa, lda) and B (n x pb, ldb).  The tile list is
// allocated from the stream-ordered pool and returned in *tiles_out for the caller to free after the stream
// has drained.
int launch_gemm_tn(cudaStream_t stream, int sm_count, const double *A, int pa, long long lda, const double *B, int pb,
                   long long ldb, long long n, double *C, long long ldc, double divisor, void **tiles_out) {
  const int nbi = (pa + BM - 1) / BM, nbj = (pb + BN_FULL - 1) / BN_FULL;
  std::vector<int2> tiles;
  tiles.reserve((size_t)nbi * nbj);
  // column panels of B outermost: the CTAs in flight share a few B panels and all of A through L2
  const int S = 8;
  for (int SJ = 0; SJ < nbj; SJ += S)
    for (int bi = 0; bi < nbi; ++bi)
      for (int bj = SJ; bj < min(SJ + S, nbj); ++bj) tiles.push_back(make_int2(bi, bj));
  const int ntiles = (int)tiles.size();
  int2 *dt = nullptr;
  CUDA_TRY(cudaMallocAsync((void **)&dt, (size_t)ntiles * sizeof(int2), stream));
  *tiles_out = dt;
  CUDA_TRY(cudaMemcpyAsync(dt, tiles.data(), (size_t)ntiles * sizeof(int2), cudaMemcpyHostToDevice, stream));
  CUDA_TRY(cudaStreamSynchronize(stream)); // `tiles` is a host temporary
  const bool aligned = ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && ((lda & 1) == 0) &&
                       ((reinterpret_cast<uintptr_t>(B) & 15) == 0) && ((ldb & 1) == 0);
  const int grid = min(ntiles, sm_count);
  const size_t dyn = (size_t)4 * (BM + BN_FULL) * (16 + 4) * sizeof(double);
  if (aligned) {
    auto kern = gram_syrk_kernel<true, 2, 4, 16, 4, true, true>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    kern<<<grid, 256, dyn, stream>>>(A, n, pa, lda, C, ldc, dt, ntiles, divisor, 1, B, pb, ldb, 0, nullptr, 0);
  } else {
    auto kern = gram_syrk_kernel<false, 2, 4, 16, 4, false, true>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    kern<<<grid, 256, dyn, stream>>>(A, n, pa, lda, C, ldc, dt, ntiles, divisor, 1, B, pb, ldb, 0, nullptr, 0);
  }
  CUDA_TRY(cudaGetLastError());
  CD_COUNT_LAUNCH(1);
  return CDGPU_OK;
}

// The same product with the ROWS split as well (few output tiles, many rows: the skinny column batches of the lazy
// covariance form): work items = tiles x row chunks, with the chunk count chosen so that the items fill whole waves
// of the SMs; partial tiles go to a slab and are summed in a fixed order (deterministic).  Everything is stream
// ordered; the scratch comes from and returns to the stream's pool.
int launch_gemm_tn_split(cudaStream_t stream, int sm_count, const double *A, int pa, long long lda, const double *B, int pb,
                         long long ldb, long long n, double *C, long long ldc, double divisor) {
  const bool aligned = ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && ((lda & 1) == 0) &&
                       ((reinterpret_cast<uintptr_t>(B) & 15) == 0) && ((ldb & 1) == 0);
  // narrow batches (<= 32 columns) use a 128 x 32 tile: a quarter of the DMMA work of the 128-wide tile for the same X traffic
  const int bn = (pb <= 32 && aligned) ? 32 : BN_FULL;
  const int nbi = (pa + BM - 1) / BM, nbj = (pb + bn - 1) / bn;
  const int ntiles = nbi * nbj;
  int best_c = 1;
  double best_eff = 0.0;
  for (int c = 1; c <= 64; ++c) {
    if (c > 1 && n / c < 512) break;
    const long long items = (long long)ntiles * c, waves = (items + sm_count - 1) / sm_count;
    const double eff = (double)items / (double)(waves * sm_count);
    if (eff > best_eff + 0.02) {
      best_eff = eff;
      best_c = c;
    }
  }
  if (!aligned || (best_c == 1 && bn == BN_FULL) || ntiles >= 65536 || nbi >= 65536) {
    void *tiles = nullptr;
    CD_TRY(launch_gemm_tn(stream, sm_count, A, pa, lda, B, pb, ldb, n, C, ldc, divisor, &tiles));
    CUDA_TRY(cudaFreeAsync(tiles, stream));
    return CDGPU_OK;
  }
  int nchunks = best_c;
  long long kchunk = ((n + nchunks - 1) / nchunks + 15) & ~15ll;
  nchunks = (int)((n + kchunk - 1) / kchunk);
  std::vector<int2> base, items;
  base.reserve((size_t)ntiles);
  for (int bj = 0; bj < nbj; ++bj)
    for (int bi = 0; bi < nbi; ++bi) base.push_back(make_int2(bi, bj));
  items.reserve((size_t)ntiles * nchunks);
  for (int ch = 0; ch < nchunks; ++ch)
    for (int t = 0; t < ntiles; ++t) items.push_back(make_int2(base[(size_t)t].x | (ch << 16), base[(size_t)t].y | (t << 16)));
  int2 *ditems = nullptr, *dbase = nullptr;
  double *slab = nullptr;
  const size_t slab_doubles = (size_t)nchunks * ntiles * BM * bn;
  CUDA_TRY(cudaMallocAsync((void **)&ditems, items.size() * sizeof(int2), stream));
  CUDA_TRY(cudaMallocAsync((void **)&dbase, base.size() * sizeof(int2), stream));
  CUDA_TRY(cudaMallocAsync((void **)&slab, slab_doubles * sizeof(double), stream));
  CUDA_TRY(cudaMemcpyAsync(ditems, items.data(), items.size() * sizeof(int2), cudaMemcpyHostToDevice, stream));
  CUDA_TRY(cudaMemcpyAsync(dbase, base.data(), base.size() * sizeof(int2), cudaMemcpyHostToDevice, stream));
  CUDA_TRY(cudaStreamSynchronize(stream)); // host temporaries
  const int grid = min((int)items.size(), sm_count);
  if (bn == 32) {
    auto kern = gram_syrk_kernel<true, 8, 1, 16, 4, true, true, 32>;
    const size_t dyn = (size_t)4 * (BM + 32) * (16 + 4) * sizeof(double);
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    kern<<<grid, 256, dyn, stream>>>(A, n, pa, lda, C, ldc, ditems, (int)items.size(), divisor, 1, B, pb, ldb, kchunk, slab, ntiles);
  } else {
    auto kern = gram_syrk_kernel<true, 2, 4, 16, 4, true, true>;
    const size_t dyn = (size_t)4 * (BM + BN_FULL) * (16 + 4) * sizeof(double);
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    kern<<<grid, 256, dyn, stream>>>(A, n, pa, lda, C, ldc, ditems, (int)items.size(), divisor, 1, B, pb, ldb, kchunk, slab, ntiles);
  }
  CUDA_TRY(cudaGetLastError());
  gemm_reduce_slabs_kernel<<<ntiles, 512, 0, stream>>>(slab, nchunks, ntiles, dbase, pa, pb, C, ldc, divisor, bn);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaFreeAsync(slab, stream));
  CUDA_TRY(cudaFreeAsync(ditems, stream));
  CUDA_TRY(cudaFreeAsync(dbase, stream));
  CD_COUNT_LAUNCH(2);
  return CDGPU_OK;
}

int launch_scale_gram(cdgpu_handle_s *h, double *G, double *c, int p, double n_total) {
  const long long ldg = ((long long)p + 1) & ~1ll;
  (void)c; // c sits right behind G in the same allocation (api.cu: gram_build)
  const long long count = ldg * p + p;
  scale_kernel<<<h->sm_count * 4, 256, 0, h->stream>>>(G, count, n_total);
  CUDA_TRY(cudaGetLastError());
  const int nb = (p + 31) / 32;
  mirror_lower_kernel<<<dim3(nb, nb), dim3(32, 8), 0, h->stream>>>(G, ldg, p);
  CUDA_TRY(cudaGetLastError());
  CD_COUNT_LAUNCH(2);
  return CDGPU_OK;
}
