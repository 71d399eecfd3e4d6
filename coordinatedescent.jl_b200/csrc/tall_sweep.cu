// Residual-form sweeps of CDSqrtLassoLoss for TALL problems (cd_differentiable_function.jl:197-291 under the driver of
// coordinate_descent.jl:7-110): n beyond what one CTA's shared memory holds (naive_sweep.cu keeps a copy of r per CTA,
// n <= ~28 000).  The reference has no such bound; the covariance form that carries tall LS / WLS problems (api.cu:
// tall_attach) does not apply to the sqrt-lasso step, which needs ||r|| at every visit.
//
// Layout: the ROWS are dealt over the CTAs of one cooperative grid (one CTA per SM); CTA g keeps rows
// [g L, (g+1) L) of r in its shared memory for the whole call (L = ceil(n / G); in global memory beyond ~25 000 rows per
// CTA, i.e. n > 3.7 M).  A visit of coordinate k needs s0 = X_k' r and ||r||^2:
//   s = s0 + x_k a_k,  rsqr+ = ||r||^2 + x_k (2 s0 + x_k a_k)      (a_k = X_k' X_k, the partial residual never formed)
// then the closed form of :277-283 and r -= X_k h.  Every CTA holds a slice partial of s0 and of ||r||^2; one grid barrier
// makes them visible, every CTA adds the G partials in the same order and takes the same decision (no broadcast).
// In a full pass a WINDOW of 64 consecutive positions is evaluated per barrier (four columns per warp against one read of
// the slice): the coordinates that do not move (almost all of a full pass) cost no barrier of their own; the first one
// that moves is applied and the window restarts behind it, so the visit sequence is the reference's.  Barriers per full
// pass = ceil(p / 64) + moves; a pass over the stored entries takes one barrier per visit (the whole CTA on one column).
// The iterate (list, values, dense copy, membership) is written by CTA 0 only; the others read x_k before the window's
// barrier and the list after the barrier that ends a pass.
#include <cooperative_groups.h>

#include "common.cuh"
namespace cg = cooperative_groups;

namespace {
constexpr int TS_T = 512, TS_NW = TS_T / 32;
constexpr int TS_W = 256, TS_PART = TS_W + 1; // most positions per window; doubles per CTA and buffer: the dots + the slice's ||r||^2
constexpr int TS_HDR = 2930;                  // shared doubles ahead of the r slice
constexpr int TS_GM = 32;                     // longest list whose passes run on its Gram (one lane per entry)
constexpr int TS_GPART = 532;                 // doubles per CTA: 32*31/2 pair dots + 32 dots with r + ... + the slice's ||r||^2 (last)

__device__ __forceinline__ double2 ts_ld2(const double2 *p) { // X is read once per visit: keep it out of L1
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
// one warp: NC column dots against the same slice of r (16-byte loads, r read once for the NC columns)
template <int NC>
__device__ __forceinline__ void ts_warp_dots(const double *const *col, const double *r, int len, int lane, double *out) {
  double acc[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) acc[c] = 0.0;
  const int pairs = len >> 1;
  const double2 *r2 = reinterpret_cast<const double2 *>(r);
#pragma unroll 2
  for (int j = lane; j < pairs; j += 32) {
    const double2 rv = r2[j];
    double2 xv[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) xv[c] = ts_ld2(reinterpret_cast<const double2 *>(col[c]) + j);
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[c] = fma(xv[c].y, rv.y, fma(xv[c].x, rv.x, acc[c]));
  }
  if ((len & 1) && lane == 0) {
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[c] = fma(__ldg(col[c] + len - 1), r[len - 1], acc[c]);
  }
#pragma unroll
  for (int c = 0; c < NC; ++c) out[c] = warp_sum(acc[c]);
}

__device__ __forceinline__ void ts_grid_sync(unsigned *ctr, unsigned &target, unsigned G) {
  target += G;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    unsigned v;
    for (;;) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if (v >= target) break;
    }
  }
  __syncthreads();
}

// descendCoordinate!(f::CDSqrtLassoLoss, ...) :242-291 from s0 = X_k' r, a = X_k' X_k, rr = ||r||^2, old = x_k
// (srr = sqrt(rr), shared by the coordinates of a window that are zero)
__device__ __forceinline__ double ts_step(double s0, double a, double old, double l, double rr, double srr) {
  const double s = old != 0.0 ? fma(old, a, s0) : s0;
  const double rsq = old != 0.0 ? rr + old * fma(old, a, 2.0 * s0) : rr;
  const double t = l * (old != 0.0 ? sqrt(rsq) : srr);
  if (fabs(s) <= t) return 0.0;
  const double q = l / sqrt(1.0 - l * l / a) * sqrt(rsq - s * s / a);
  return s > t ? (s - q) / a : (s + q) / a;
}

__global__ void __launch_bounds__(TS_T, 1) tall_sqrt_kernel(const NaiveArgs a, double *part, int L, int r_in_smem, int wmax, double *gpart) {
  extern __shared__ __align__(16) double ts_sm[];
  double *red = ts_sm;           // [0, TS_PART): sums over the CTAs of the window's dots and of ||r||^2
  double *wred = ts_sm + 264;    // [264, 280): per-warp partials of a block reduction
  double *slice_rr = ts_sm + 280; // this CTA's sum of r_i^2
  double *tmp1 = ts_sm + 281;
  int *smv = reinterpret_cast<int *>(ts_sm + 282); // [282, 290): per warp of deciders, first mover [0, 8) / first stored entry [8, 16)
  int *sk = reinterpret_cast<int *>(ts_sm + 290);  // [290, 418): the window's coordinates
  // passes over the stored entries on the Gram of the list
  double *Gs = ts_sm + 1700, *ds = ts_sm + 2724, *bs = ts_sm + 2756, *b0 = ts_sm + 2788, *acol = ts_sm + 2820, *lamv = ts_sm + 2852;
  int *ord = reinterpret_cast<int *>(ts_sm + 2884), *crd = reinterpret_cast<int *>(ts_sm + 2900);
  double *res = ts_sm + 2916; // maxH, ||r||^2, visits
  int *resi = reinterpret_cast<int *>(ts_sm + 2920); // list length, passes, converged, accepted
  double *sxk = ts_sm + 418, *scol = ts_sm + 674, *slam = ts_sm + 930, *sh = ts_sm + 1186, *snw = ts_sm + 1442; // x_k, a_k, omega_k, h, new x_k
  const int G = gridDim.x, bid = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long row0 = (long long)bid * L;
  const int len = (int)max(0ll, min((long long)L, (long long)a.n - row0));
  double *r = r_in_smem ? ts_sm + TS_HDR : a.r + row0;
  // 16-byte loads of the column slices (row0 is even): X 16-byte aligned with an even leading dimension
  const bool vec = (a.ldx & 1) == 0 && (reinterpret_cast<unsigned long long>(a.X) & 15) == 0 &&
                   (reinterpret_cast<unsigned long long>(a.r) & 15) == 0;
  unsigned *ctr = reinterpret_cast<unsigned *>(a.flag + 7);
  unsigned target = 0;
  if (bid == 0 && tid == 0) __stcg(ctr, 0u);
  if (r_in_smem)
    for (int i = tid; i < len; i += TS_T) r[i] = a.r[row0 + i];
  __syncthreads();
  cg::this_grid().sync(); // the barrier counter is zero before anyone arrives

  // block sum of v -> *slice_rr (all threads may read it after the call)
  auto block_sum_to = [&](double v, double *dst) {
    v = warp_sum(v);
    if (lane == 0) wred[warp] = v;
    __syncthreads();
    if (warp == 0) {
      double t = lane < TS_NW ? wred[lane] : 0.0;
      t = warp_sum(t);
      if (lane == 0) *dst = t;
    }
    __syncthreads();
  };
  {
    double acc = 0.0;
    for (int i = tid; i < len; i += TS_T) acc = fma(r[i], r[i], acc);
    block_sum_to(acc, slice_rr);
  }

  int nact = *a.nact; // every CTA's view of the list length (refreshed behind the barrier that ends a pass)
  int buf = 0;
  unsigned long long pass_counter = 0;
  DevStats st;
  st.passes = st.full_passes = st.visits = st.accepted = 0;
  st.maxH = 0.0;
  st.converged = 0;
  st.outer_iters = 0;
  st.sigma = 0.0;
  int status = 0;
  long long cols_done = 0, out_off = 0;
  const bool ordered = a.randomize == 0;

  for (int li = 0; li < a.nlambda && status == 0; ++li) {
    if (li > 0 && !a.accumulate) {
      st.passes = st.full_passes = st.visits = st.accepted = 0;
      st.maxH = 0.0;
    }
    const double lam = a.lambdas[li];
    // ---- _coordinateDescent! (coordinate_descent.jl:65-92)
    st.converged = 0;
    bool conv = true;
    for (long long iter = 1; iter <= a.maxIter; ++iter) {
      if (!conv && gpart && nact >= 1 && nact <= TS_GM) {
        // ---- the passes over the stored entries (until one of them converges) on the Gram of the list: with
        // G = X_A' X_A and d = X_A' r a visit is s0 = d_g, and a step h of entry g is d -= h G[:, g],
        // ||r||^2 += h (h a_g - 2 s0); r itself is brought up to date once at the end.  Two grid barriers for the whole
        // phase instead of one per visit; every CTA runs the same chain on the same sums (warp 0, one lane per entry).
        const int m0 = nact, npair = m0 * (m0 - 1) / 2, nval = npair + m0;
        if (tid < m0) {
          const int k = __ldcg(a.act + tid);
          crd[tid] = k;
          const double b = __ldcg(a.beta + k);
          b0[tid] = b;
          bs[tid] = b;
          acol[tid] = __ldg(a.colsq + k);
          lamv[tid] = a.omega ? __ldg(a.omega + k) : 1.0;
          ord[tid] = tid;
        }
        __syncthreads();
        double *gmine = gpart + (long long)bid * TS_GPART;
        for (int q = warp; q < nval; q += TS_NW) {
          int i = 0, j;
          const double *y;
          if (q < npair) { // pair (i, j), i < j
            int rem = q;
            while (rem >= m0 - 1 - i) {
              rem -= m0 - 1 - i;
              i += 1;
            }
            j = i + 1 + rem;
            y = a.X + (long long)crd[j] * a.ldx + row0;
          } else {
            i = q - npair;
            y = r;
          }
          const double *x = a.X + (long long)crd[i] * a.ldx + row0;
          double acc0 = 0.0, acc1 = 0.0;
          int t = lane;
          for (; t + 32 < len; t += 64) {
            acc0 = fma(__ldg(x + t), y[t], acc0);
            acc1 = fma(__ldg(x + t + 32), y[t + 32], acc1);
          }
          if (t < len) acc0 = fma(__ldg(x + t), y[t], acc0);
          const double v = warp_sum(acc0 + acc1);
          if (lane == 0) __stcg(gmine + q, v);
        }
        if (tid == 0) __stcg(gmine + TS_GPART - 1, *slice_rr);
        ts_grid_sync(ctr, target, (unsigned)G);
        for (int q = warp; q <= nval; q += TS_NW) {
          const double *src = gpart + (q < nval ? q : TS_GPART - 1);
          double t = 0.0;
          for (int g = lane; g < G; g += 32) t += __ldcg(src + (long long)g * TS_GPART);
          t = warp_sum(t);
          if (lane == 0) {
            if (q < npair) {
              int i = 0, rem = q;
              while (rem >= m0 - 1 - i) {
                rem -= m0 - 1 - i;
                i += 1;
              }
              const int j = i + 1 + rem;
              Gs[i * TS_GM + j] = t;
              Gs[j * TS_GM + i] = t;
            } else if (q < nval) {
              ds[q - npair] = t;
            } else {
              res[1] = t;
            }
          }
        }
        if (tid < m0) Gs[tid * TS_GM + tid] = acol[tid];
        __syncthreads();
        if (warp == 0) {
          const bool on = lane < m0;
          double d = on ? ds[lane] : 0.0, be = on ? bs[lane] : 0.0;
          const double ac = on ? acol[lane] : 1.0, lm = on ? lamv[lane] * lam : 0.0;
          double rr = res[1], maxH = 0.0;
          int mcur = m0, acc = 0;
          long long np = 0, vis = 0;
          const long long left = a.maxIter - iter + 1;
          bool cv = false;
          while (np < left) {
            const PermKey pk = cd_perm_key((uint32_t)max(mcur, 1), a.seed, pass_counter + (unsigned long long)np);
            maxH = 0.0;
            for (int i = 0; i < mcur; ++i) {
              const int g = ord[ordered ? i : (int)cd_perm(pk, (uint32_t)i)];
              const double s0 = __shfl_sync(0xffffffffu, d, g), old = __shfl_sync(0xffffffffu, be, g);
              const double ag = __shfl_sync(0xffffffffu, ac, g), lg = __shfl_sync(0xffffffffu, lm, g);
              const double nw = ts_step(s0, ag, old, lg, rr, sqrt(rr));
              const double h = nw - old;
              if (fabs(h) > maxH) maxH = fabs(h);
              if (h != 0.0) {
                acc += 1;
                if (on) d = fma(-h, Gs[lane * TS_GM + g], d);
                rr = fma(h, fma(h, ag, -2.0 * s0), rr);
                if (lane == g) be = nw;
              }
            }
            vis += mcur;
            np += 1;
            if (on) bs[lane] = be;
            __syncwarp();
            if (lane == 0) { // dropzeros!: the last stored entry moves into a hole
              int i = 0, mm = mcur;
              while (i < mm) {
                if (bs[ord[i]] == 0.0) {
                  if (i != mm - 1) ord[i] = ord[mm - 1];
                  mm -= 1;
                } else {
                  i += 1;
                }
              }
              resi[0] = mm;
            }
            __syncwarp();
            mcur = resi[0];
            cv = maxH < a.optTol;
            if (cv) break;
          }
          if (lane == 0) {
            res[0] = maxH;
            res[2] = (double)vis;
            resi[1] = (int)np;
            resi[2] = cv ? 1 : 0;
            resi[3] = acc;
          }
        }
        __syncthreads();
        const int mcur = resi[0], np = resi[1];
        { // r -= X_A (x - x at entry), ||r||^2 of the slice afresh (a thread owns the same rows for every column)
          for (int g = 0; g < m0; ++g) {
            const double dlt = bs[g] - b0[g];
            if (dlt == 0.0) continue;
            const double *col = a.X + (long long)crd[g] * a.ldx + row0;
            for (int i = tid; i < len; i += TS_T) r[i] = fma(-__ldg(col + i), dlt, r[i]);
          }
          double acc = 0.0;
          for (int i = tid; i < len; i += TS_T) acc = fma(r[i], r[i], acc);
          block_sum_to(acc, slice_rr);
        }
        if (bid == 0) {
          if (tid < m0) {
            __stcg(a.beta + crd[tid], bs[tid]);
            a.inlist[crd[tid]] = 0;
          }
          __syncthreads();
          if (tid < mcur) {
            const int g = ord[tid];
            a.act[tid] = crd[g];
            a.actval[tid] = bs[g];
            a.inlist[crd[g]] = 1;
          }
          if (tid == 0) *a.nact = mcur;
        }
        ts_grid_sync(ctr, target, (unsigned)G);
        nact = mcur;
        st.passes += np;
        st.visits += (long long)res[2];
        st.accepted += resi[3];
        st.maxH = res[0];
        pass_counter += (unsigned long long)np;
        conv = resi[2] != 0;
        iter += np - 1;
        __syncthreads(); // the phase's shared results are read; the next pass may refill the header
        continue;
      }
      const bool full = conv;
      const int seqlen = full ? a.p : nact;
      const PermKey pk = cd_perm_key((uint32_t)max(seqlen, 1), a.seed, pass_counter);
      pass_counter += 1;
      st.passes += 1;
      if (full) st.full_passes += 1;
      double maxH = 0.0;
      int nlist = nact; // CTA 0, thread 0: the list grows during a full pass
      // ---- _cdPass! (:94-110), a window of TS_W positions per barrier
      int i0 = 0, wcap = wmax;
      while (i0 < seqlen) {
        // a sparse pass moves at almost every visit: one position per barrier, the whole CTA on its column;
        // a full pass moves rarely: up to TS_W positions per barrier, up to four columns per warp.  What lies behind a
        // mover in its window was read for nothing, so a window ends at the first stored entry (x_k != 0: it will move)
        // and is narrower (wcap) right after a coordinate has entered
        int wn = full ? min(wcap, seqlen - i0) : 1;
        if (tid < TS_W) {
          double xk = 0.0;
          if (tid < wn) {
            const int pos = ordered ? i0 + tid : (int)cd_perm(pk, (uint32_t)(i0 + tid));
            const int k = full ? pos : __ldcg(a.act + pos);
            sk[tid] = k;
            xk = __ldcg(a.beta + k); // read before the barrier: CTA 0 writes x_k only behind it
            sxk[tid] = xk;
            scol[tid] = __ldg(a.colsq + k);
            slam[tid] = a.omega ? __ldg(a.omega + k) : 1.0;
          }
          const unsigned bal = __ballot_sync(0xffffffffu, xk != 0.0);
          if (lane == 0) smv[8 + warp] = bal ? __ffs(bal) + 32 * warp : 1 << 20;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < TS_W / 32; ++q) wn = min(wn, smv[8 + q]);
        double *mine = part + ((long long)buf * G + bid) * TS_PART;
        if (wn == 1) {
          const double *col = a.X + (long long)sk[0] * a.ldx + row0;
          double acc0 = 0.0, acc1 = 0.0;
          if (vec) {
            const int pairs = len >> 1;
            const double2 *c2 = reinterpret_cast<const double2 *>(col), *r2 = reinterpret_cast<const double2 *>(r);
            int j = tid;
            for (; j + TS_T < pairs; j += 2 * TS_T) {
              const double2 x0 = ts_ld2(c2 + j), x1 = ts_ld2(c2 + j + TS_T), r0 = r2[j], r1 = r2[j + TS_T];
              acc0 = fma(x0.y, r0.y, fma(x0.x, r0.x, acc0));
              acc1 = fma(x1.y, r1.y, fma(x1.x, r1.x, acc1));
            }
            if (j < pairs) {
              const double2 x0 = ts_ld2(c2 + j), r0 = r2[j];
              acc0 = fma(x0.y, r0.y, fma(x0.x, r0.x, acc0));
            }
            if ((len & 1) && tid == 0) acc1 = fma(__ldg(col + len - 1), r[len - 1], acc1);
          } else {
            for (int i = tid; i < len; i += TS_T) acc0 = fma(__ldg(col + i), r[i], acc0);
          }
          block_sum_to(acc0 + acc1, tmp1);
          if (tid == 0) __stcg(mine, *tmp1);
        } else if (warp < wn) {
          const int nc = (wn - warp + TS_NW - 1) / TS_NW; // positions warp, warp + 16, ...: four at a time
          for (int c0 = 0; c0 < nc; c0 += 4) {
            const int n4 = min(4, nc - c0);
            const double *col[4];
            double d[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) col[c] = a.X + (long long)sk[min(warp + (c0 + c) * TS_NW, wn - 1)] * a.ldx + row0;
            if (vec) {
              if (n4 == 4) ts_warp_dots<4>(col, r, len, lane, d);
              else if (n4 == 3) ts_warp_dots<3>(col, r, len, lane, d);
              else if (n4 == 2) ts_warp_dots<2>(col, r, len, lane, d);
              else ts_warp_dots<1>(col, r, len, lane, d);
            } else {
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                double acc = 0.0;
                if (c < n4)
                  for (int i = lane; i < len; i += 32) acc = fma(__ldg(col[c] + i), r[i], acc);
                d[c] = warp_sum(acc);
              }
            }
            if (lane == 0) {
#pragma unroll
              for (int c = 0; c < 4; ++c)
                if (c < n4) __stcg(mine + warp + (c0 + c) * TS_NW, d[c]);
            }
          }
        }
        if (tid == 0) __stcg(mine + TS_W, *slice_rr);
        ts_grid_sync(ctr, target, (unsigned)G);
        // sums over the CTAs, the same order everywhere: values v = warp, warp + 16, ... (warp 0 also the ||r||^2 slot)
        for (int c0 = 0; c0 <= TS_W / TS_NW; c0 += 4) {
          double t[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int v = warp + (c0 + c) * TS_NW;
            t[c] = 0.0;
            if (v < TS_PART && (v < wn || v == TS_W)) {
              const double *src = part + (long long)buf * G * TS_PART + v;
              for (int g = lane; g < G; g += 32) t[c] += __ldcg(src + (long long)g * TS_PART);
            }
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int v = warp + (c0 + c) * TS_NW;
            if (v < TS_PART && (v < wn || v == TS_W)) {
              const double tv = warp_sum(t[c]);
              if (lane == 0) red[v] = tv;
            }
          }
          if (c0 * TS_NW >= wn && c0 + 4 <= TS_W / TS_NW) c0 = TS_W / TS_NW - 4; // nothing left but the ||r||^2 slot
        }
        buf ^= 1;
        __syncthreads();
        // the coordinates ahead of the first one that moves leave r as it is, so every position of the window is decided
        // on its own (thread w: position w) and the first mover ends the window
        const double rr = red[TS_W];
        if (tid < TS_W) {
          double nw = 0.0, h = 0.0;
          if (tid < wn) {
            nw = ts_step(red[tid], scol[tid], sxk[tid], slam[tid] * lam, rr, sqrt(rr));
            h = nw - sxk[tid];
            sh[tid] = h;
            snw[tid] = nw;
          }
          const unsigned bal = __ballot_sync(0xffffffffu, h != 0.0);
          if (lane == 0) smv[warp] = bal ? __ffs(bal) - 1 + 32 * warp : -1;
        }
        __syncthreads();
        int moved = -1;
#pragma unroll
        for (int q = TS_W / 32 - 1; q >= 0; --q) moved = smv[q] >= 0 ? smv[q] : moved;
        double hmv = 0.0, nwv = 0.0;
        if (moved >= 0) {
          hmv = sh[moved];
          nwv = snw[moved];
          st.visits += moved + 1;
          st.accepted += 1;
          if (fabs(hmv) > maxH) maxH = fabs(hmv);
        } else {
          st.visits += wn;
        }
        if (moved >= 0) {
          const int k = sk[moved];
          const double *col = a.X + (long long)k * a.ldx + row0;
          double acc = 0.0;
          if (vec) {
            const int pairs = len >> 1;
            const double2 *c2 = reinterpret_cast<const double2 *>(col);
            double2 *r2 = reinterpret_cast<double2 *>(r);
            for (int j = tid; j < pairs; j += TS_T) {
              const double2 x = ts_ld2(c2 + j);
              double2 v = r2[j];
              v.x = fma(-x.x, hmv, v.x);
              v.y = fma(-x.y, hmv, v.y);
              r2[j] = v;
              acc = fma(v.y, v.y, fma(v.x, v.x, acc));
            }
            if ((len & 1) && tid == 0) {
              const double v = fma(-__ldg(col + len - 1), hmv, r[len - 1]);
              r[len - 1] = v;
              acc = fma(v, v, acc);
            }
          } else {
            for (int i = tid; i < len; i += TS_T) {
              const double v = fma(-__ldg(col + i), hmv, r[i]);
              r[i] = v;
              acc = fma(v, v, acc);
            }
          }
          if (bid == 0 && tid == 0) { // x[k] = newVal: appended on the first non-zero store (setindex!)
            __stcg(a.beta + k, nwv);
            if (nwv != 0.0 && !a.inlist[k]) {
              a.inlist[k] = 1;
              a.act[nlist] = k;
              nlist += 1;
            }
          }
          block_sum_to(acc, slice_rr); // (its barriers also order the window's shared arrays against the next fill)
          i0 += moved + 1;
          if (sxk[moved] == 0.0) wcap = TS_NW; // an entering coordinate: others tend to follow closely
        } else {
          __syncthreads();
          i0 += wn;
          wcap = min(wmax, 2 * wcap);
        }
      }
      // ---- dropzeros!(x) on CTA 0 (the last stored entry moves into a hole), then everyone learns the new list
      if (bid == 0) {
        __syncthreads();
        if (tid == 0) sk[0] = nlist;
        __syncthreads();
        const int m = sk[0];
        for (int i = tid; i < m; i += TS_T) a.actval[i] = __ldcg(a.beta + a.act[i]);
        __syncthreads();
        if (tid == 0) {
          int mm = m, i = 0;
          while (i < mm) {
            if (a.actval[i] == 0.0) {
              a.inlist[a.act[i]] = 0;
              if (i != mm - 1) {
                a.actval[i] = a.actval[mm - 1];
                a.act[i] = a.act[mm - 1];
              }
              mm -= 1;
            } else {
              i += 1;
            }
          }
          *a.nact = mm;
        }
      }
      ts_grid_sync(ctr, target, (unsigned)G);
      nact = __ldcg(a.nact);
      st.maxH = maxH;
      const bool prev = conv;
      conv = maxH < a.optTol;
      if (prev && conv) {
        st.converged = 1;
        break;
      }
    }
    // ---- end of this lambda
    const int nnz = nact;
    if (!a.accumulate) {
      if (a.colptr && out_off + nnz > a.capacity) status = 1;
      if (bid == 0 && status == 0) {
        if (a.colptr) {
          for (int i = tid; i < nnz; i += TS_T) {
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
  if (r_in_smem)
    for (int i = tid; i < len; i += TS_T) a.r[row0 + i] = r[i];
  if (bid == 0 && tid == 0) {
    if (a.accumulate && a.stats) a.stats[0] = st;
    a.flag[0] = status;
    a.flag[1] = (int)cols_done;
  }
}
} // namespace

// tall CDSqrtLassoLoss handles (and any sqrt-lasso handle under CDGPU_FORCE_TALL=1): same contract as launch_naive_path
int launch_tall_sqrt(cdgpu_handle_s *h, const NaiveArgs &a) {
  if (a.kind != CDGPU_LOSS_SQRT || a.scaled)
    return cdgpu_set_error(CDGPU_EARG, "the row-distributed sweep is the sqrt-lasso one");
  static bool attr_done[64] = {false};
  const size_t max_dyn = 227 * 1024;
  const bool known = h->device >= 0 && h->device < 64;
  if (!known || !attr_done[h->device]) {
    CUDA_TRY(cudaFuncSetAttribute(tall_sqrt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_dyn));
    if (known) attr_done[h->device] = true;
  }
  int G = h->sm_count;
  if (const char *env = getenv("CDGPU_TALL_GRID")) G = std::max(1, std::min(G, atoi(env)));
  G = (int)std::min<long long>(G, ((long long)a.n + 63) / 64); // at least 64 rows per CTA
  const int L = (int)((((long long)a.n + G - 1) / G + 1) & ~1ll);
  size_t dyn = (size_t)(TS_HDR + L) * sizeof(double);
  int r_in_smem = 1;
  if (dyn > max_dyn || getenv("CDGPU_TALL_R_GLOBAL")) {
    r_in_smem = 0;
    dyn = (size_t)TS_HDR * sizeof(double);
  }
  const size_t part_doubles = (size_t)2 * h->sm_count * TS_PART + (size_t)h->sm_count * TS_GPART;
  if (!h->dtall) CUDA_TRY(cudaMalloc((void **)&h->dtall, part_doubles * sizeof(double)));
  double *part = h->dtall;
  int Larg = L;
  // positions per window of a full pass: about 400 MB of columns per barrier, 64..256
  int wmax = (int)std::min<long long>(TS_W, std::max<long long>(64, (400000000ll / (8ll * a.n) + 63) / 64 * 64));
  if (const char *env = getenv("CDGPU_TALL_WINDOW")) wmax = std::max(1, std::min(TS_W, atoi(env)));
  double *gpart = part + (size_t)2 * h->sm_count * TS_PART;
  if (const char *env = getenv("CDGPU_TALL_GRAM"))
    if (atoi(env) == 0) gpart = nullptr; // passes over the stored entries visit by visit
  void *args[] = {(void *)&a, (void *)&part, (void *)&Larg, (void *)&r_in_smem, (void *)&wmax, (void *)&gpart};
  CUDA_TRY(cudaLaunchCooperativeKernel(tall_sqrt_kernel, dim3(G), dim3(TS_T), args, dyn, h->stream));
  CD_COUNT_LAUNCH(1);
  return CDGPU_OK;
}
