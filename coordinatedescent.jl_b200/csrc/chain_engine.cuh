// chain_engine.cuh — the sequential Gauss-Seidel chain of ACTIVE-SET passes (_cdPass! over the stored
// entries, src/coordinate_descent.jl:94-110 with the sparse iterators of src/atom_iterator.jl:18-37,
// 53-75) as a blocked, software-pipelined CTA-level engine.  Shared by the covariance-form kernel
// (cov_sweep.cu) and by the naive kernels once they have formed the active Gram (naive_sweep.cu).
//
// State per stored entry t (list position): g_t (cov: (A x)_t; naive: X_t'(w.r)), beta_t.  A step on
// entry i changes every g_t by G[t,i]*h, so a pass is a length-m dependent chain.  Layout of the work:
//   * visit positions are cut into blocks of 32.  WARP 0 runs the chain of one block entirely in
//     registers: lane j owns entry j of the block, a step is "all lanes evaluate their closed-form
//     update, lane i's h is broadcast by one shuffle, every lane applies G[j,i]*h" — no barrier and
//     no memory access on the dependent path (the 32x32 diagonal block D sits in shared memory and
//     its next row is fetched ahead of the shuffle).
//   * before the chain of block b, warp 0 applies the 32 steps of block b-1 to its new entries (the
//     "panel" P = G[block b, block b-1], also staged in shared memory).
//   * WARPS 1..15 meanwhile (a) gather D, P and the per-entry constants of block b+1 from L2 into the
//     other stage buffer and (b) apply the steps of block b-1 to every entry outside blocks b-1 and b.
//     One __syncthreads per block hands the results over.
//   * every entry receives the steps in visit order, each as the same non-fused multiply-add the
//     sequential algorithm performs, so iterates are bit-identical to the one-step-at-a-time chain.
//   * dropzeros! (swap-with-last compaction, ProximalBase) runs at the end of a pass, only when an
//     entry became exactly zero.
#pragma once
#include "common.cuh"

namespace chain {

#ifdef CHAIN_PROBE
__device__ int probe_mode; // benchmarks/micro/chain_probe.cu: bit 0 skip worker apply, bit 1 skip worker stage, bit 2 skip chain
#define CHAIN_PROBE_BIT(b) ((probe_mode >> (b)) & 1)
#else
#define CHAIN_PROBE_BIT(b) 0
#endif

constexpr int BUF_DOUBLES = 2 * 32 * 32 + 3 * 32; // D, P, three per-entry constants
constexpr int STAGE_DOUBLES = 2 * BUF_DOUBLES;

struct Shared { // small block-shared scalars
  double hb[2][32];
  double pmax;
  int newm, flag;
};

struct State {
  int m;                       // stored entries
  int *row;                    // [cap] row/column id of the entry inside G
  int *coord;                  // [cap] coordinate (0-based) of the entry (may alias row)
  double *g, *be;              // [cap]
  unsigned short *ord, *pos;   // [cap] visit position -> entry, entry -> visit position
  double *stage;               // [STAGE_DOUBLES]
  Shared *sh;
  const double *G;             // symmetric; G(t,i) = G[row[t] + row[i]*ldg]
  long long ldg;
  long long *prof; // optional [8]: cycles of thread 0 (panel, chain, barrier wait, pass ends) and thread 32 (stage, apply, barrier wait), blocks
  const int *slot = nullptr; // optional: column r of G lives at G + slot[r]*ldg (lazily formed covariance columns); null: r
  double *hout = nullptr; // optional [cap]: ONE-PASS mode (run() only, ordered, maxPasses = 1): the step h of every entry is
                          // recorded here and dropzeros! is left to the caller (the member chain of a FULL pass, cov_sweep.cu)
};

struct Result {
  long long npasses, visits, accepted;
  double maxH;
  int conv, m;
};

__device__ __forceinline__ double ld_l2(const double *p) { return __ldcg(p); }
// start of column r of G
__device__ __forceinline__ int gcolidx(const State &S, int r) { return S.slot ? __ldg(S.slot + r) : r; }
__device__ __forceinline__ const double *gcol(const State &S, int r) { return S.G + (long long)gcolidx(S, r) * S.ldg; }

// gather block `bb` (diagonal block, panel against block bb-1, constants) into `buf`; threads t0, t0+nthr, ...
template <class Policy>
__device__ __forceinline__ void stage_block(const State &S, const Policy &P, int m, int bb, double *buf, int t0, int nthr) {
  const int cnt = min(32, m - 32 * bb);
  const int total = bb > 0 ? 2048 : 1024;
  for (int base = t0; base < total; base += 5 * nthr) {
    double v[5];
#pragma unroll
    for (int u = 0; u < 5; ++u) {
      const int idx = base + u * nthr;
      v[u] = 0.0;
      if (idx < total) {
        const int i = (idx >> 5) & 31, j = idx & 31;
        const bool panel = idx >= 1024;
        if (j < cnt && (panel || i < cnt)) {
          const int rj = S.row[S.ord[32 * bb + j]];
          const int ri = S.row[S.ord[32 * (panel ? bb - 1 : bb) + i]];
          v[u] = ld_l2(gcol(S, ri) + rj);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 5; ++u) {
      const int idx = base + u * nthr;
      if (idx < total) buf[idx] = v[u];
    }
  }
  for (int j = t0; j < cnt; j += nthr) {
    double c0, c1, c2;
    P.load_consts(S.coord[S.ord[32 * bb + j]], c0, c1, c2);
    buf[2048 + j] = c0;
    buf[2048 + 32 + j] = c1;
    buf[2048 + 64 + j] = c2;
  }
}

// apply the steps of block `hbk` (h values in hv[0..cnt)) to every entry whose block is neither ex0 nor ex1
template <class Policy>
__device__ __forceinline__ void apply_block(const State &S, int m, int hbk, const double *hv, int ex0, int ex1, int t0, int nthr) {
  const int cnt = min(32, m - 32 * hbk);
  bool any = false;
  for (int i = 0; i < cnt; ++i) any |= hv[i] != 0.0;
  if (!any) return;
  for (int t = t0; t < m; t += nthr) {
    const int blk = S.pos[t] >> 5;
    if (blk == ex0 || blk == ex1) continue;
    const int rt = S.row[t];
    double gt = S.g[t];
#pragma unroll 1
    for (int i0 = 0; i0 < cnt; i0 += 16) {
      double v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u)
        v[u] = (i0 + u < cnt) ? ld_l2(gcol(S, S.row[S.ord[32 * hbk + i0 + u]]) + rt) : 0.0;
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const double h = (i0 + u < cnt) ? hv[i0 + u] : 0.0;
        if (h != 0.0) gt = Policy::apply(gt, v[u], h);
      }
    }
    S.g[t] = gt;
  }
}

// Runs consecutive active-set passes until one has max|h| < optTol or maxPasses are used.
// Block-collective over T threads (T >= 64, multiple of 32); S.m entries are loaded in S.g/be/row/coord.
template <int T, class Policy>
__device__ Result run(State &S, const Policy &P, double rr, long long maxPasses, unsigned long long pass_counter,
                      bool ordered, unsigned long long seed, double optTol, unsigned char *inlist) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NWORK = T - 32;
  int m = S.m;
  Result R{0, 0, 0, 0.0, 0, m};
  Shared *sh = S.sh;
  long long pc[4] = {0, 0, 0, 0};
  const bool prof = S.prof != nullptr && (tid == 0 || tid == 32);
  long long tp = prof ? clock64() : 0;
  auto lap = [&](int slot) {
    if (prof) {
      const long long t = clock64();
      pc[slot] += t - tp;
      tp = t;
    }
  };
  bool staged_ok = false; // ordered, <= 2 blocks: both stage buffers still hold this list's blocks
  while (R.npasses < maxPasses) {
    const PermKey pkm = cd_perm_key((uint32_t)max(m, 1), seed, pass_counter + R.npasses);
    for (int s = tid; s < m; s += T) {
      const int e = ordered ? s : (int)cd_perm(pkm, (uint32_t)s);
      S.ord[s] = (unsigned short)e;
      S.pos[e] = (unsigned short)s;
    }
    __syncthreads();
    const int nb = (m + 31) >> 5;
    if (!staged_ok && nb > 0) {
      stage_block(S, P, m, 0, S.stage, tid, T);
      __syncthreads();
    }
    double pmax = 0.0;
    long long acc = 0;
    lap(3);
    for (int b = 0; b < nb; ++b) {
      double *buf = S.stage + (b & 1) * BUF_DOUBLES;
      if (warp == 0) {
        const int cnt = min(32, m - 32 * b);
        const bool valid = lane < cnt;
        const int e = valid ? S.ord[32 * b + lane] : 0;
        double gj = valid ? S.g[e] : 0.0, bej = valid ? S.be[e] : 0.0;
        // lanes past the end of the list step on benign constants (their results are never used; garbage could send
        // every step of the warp through a policy's slow path)
        const double c0 = valid ? buf[2048 + lane] : 1.0, c1 = valid ? buf[2048 + 32 + lane] : 1.0, c2 = valid ? buf[2048 + 64 + lane] : 1.0;
        if (b > 0) { // steps of the previous block, in order
          const double *hp = sh->hb[(b - 1) & 1], *Pb = buf + 1024;
#pragma unroll 1
          for (int i0 = 0; i0 < 32; i0 += 16) { // loads first, then the dependent adds
            double pv[16], hv[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
              pv[u] = Pb[(i0 + u) * 32 + lane];
              hv[u] = hp[i0 + u];
            }
#pragma unroll
            for (int u = 0; u < 16; ++u)
              if (hv[u] != 0.0) gj = Policy::apply(gj, pv[u], hv[u]);
          }
        }
        lap(0);
        double myh = 0.0;
        double drow = buf[lane];
        for (int i = 0; i < (CHAIN_PROBE_BIT(2) ? 0 : cnt); ++i) {
          const double dnext = buf[((i + 1) & 31) * 32 + lane];
          double nw, hi, dr;
          P.step(gj, bej, c0, c1, c2, rr, nw, hi, dr);
          const double h = __shfl_sync(0xffffffffu, hi, i);
          if (lane == i) {
            bej = nw;
            myh = hi;
          }
          if (h != 0.0) {
            gj = Policy::apply(gj, drow, h);
            if (Policy::HAS_RR) rr += __shfl_sync(0xffffffffu, dr, i);
            acc += 1;
          }
          drow = dnext;
        }
        if (valid) {
          S.g[e] = gj;
          S.be[e] = bej;
          if (S.hout) S.hout[e] = myh;
        }
        sh->hb[b & 1][lane] = valid ? myh : 0.0;
        pmax = fmax(pmax, fabs(myh));
        lap(1);
      } else {
        const int wt = tid - 32;
        if (b + 1 < nb && !staged_ok && !CHAIN_PROBE_BIT(1)) stage_block(S, P, m, b + 1, S.stage + ((b + 1) & 1) * BUF_DOUBLES, wt, NWORK);
        lap(0);
        if (b >= 1 && !CHAIN_PROBE_BIT(0)) apply_block<Policy>(S, m, b - 1, sh->hb[(b - 1) & 1], b - 1, b, wt, NWORK);
        lap(1);
      }
      __syncthreads();
      lap(2);
    }
    if (S.hout) {
      // one-pass mode: the caller only wants the steps and the new values; g is recomputed from the slices
    } else if (nb == 2) {
      // drain, two blocks: block 1's steps reach block 0 through the panel already staged for block 1
      // (P[i*32 + j] = G(entry j of block 1, entry i of block 0), G symmetric) — no trip to L2
      if (warp == 1) {
        const int cnt1 = m - 32;
        const double *Pb = S.stage + BUF_DOUBLES + 1024, *hv = sh->hb[1];
        const int e = S.ord[lane];
        double gt = S.g[e];
        for (int j = 0; j < cnt1; ++j) {
          const double h = hv[j];
          if (h != 0.0) gt = Policy::apply(gt, Pb[lane * 32 + j], h);
        }
        S.g[e] = gt;
      }
    } else if (nb > 2) { // drain: the last block's steps reach the rest of the list
      apply_block<Policy>(S, m, nb - 1, sh->hb[(nb - 1) & 1], nb - 1, -1, tid, T);
    }
    if (warp == 0) {
      pmax = warp_max(pmax);
      if (lane == 0) sh->pmax = pmax;
    }
    // ---- dropzeros!
    int z = 0;
    if (!S.hout)
      for (int i = tid; i < m; i += T) z |= (S.be[i] == 0.0);
    z = __syncthreads_or(z);
    R.npasses += 1;
    R.visits += m;
    R.accepted += __shfl_sync(0xffffffffu, acc, 0); // only meaningful on warp 0; fixed up below
    R.maxH = sh->pmax;
    staged_ok = ordered && nb <= 2 && !z;
    if (z) { // rare: an entry left the active set — swap-with-last in list order
      unsigned short *idx = S.ord;
      for (int i = tid; i < m; i += T) idx[i] = (unsigned short)i;
      __syncthreads();
      if (tid == 0) {
        int n = m, i = 0;
        while (i < n) {
          if (S.be[idx[i]] == 0.0) {
            inlist[S.coord[idx[i]]] = 0;
            if (i != n - 1) idx[i] = idx[n - 1];
            n -= 1;
          } else {
            i += 1;
          }
        }
        sh->newm = n;
      }
      __syncthreads();
      const int mn = sh->newm;
      double *tmpd = S.stage;
      int *tmpi = reinterpret_cast<int *>(S.stage);
      // permute the four per-entry arrays through the stage area (>= 4096 doubles)
      for (int i = tid; i < mn; i += T) tmpd[i] = S.g[idx[i]];
      __syncthreads();
      for (int i = tid; i < mn; i += T) S.g[i] = tmpd[i];
      __syncthreads();
      for (int i = tid; i < mn; i += T) tmpd[i] = S.be[idx[i]];
      __syncthreads();
      for (int i = tid; i < mn; i += T) S.be[i] = tmpd[i];
      __syncthreads();
      for (int i = tid; i < mn; i += T) tmpi[i] = S.row[idx[i]];
      __syncthreads();
      for (int i = tid; i < mn; i += T) S.row[i] = tmpi[i];
      __syncthreads();
      if (S.coord != S.row) {
        for (int i = tid; i < mn; i += T) tmpi[i] = S.coord[idx[i]];
        __syncthreads();
        for (int i = tid; i < mn; i += T) S.coord[i] = tmpi[i];
        __syncthreads();
      }
      m = mn;
    }
    if (R.maxH < optTol) {
      R.conv = 1;
      break;
    }
  }
  lap(3);
  if (prof) {
    const int o = tid == 0 ? 0 : 4;
    for (int i = 0; i < 4; ++i) S.prof[o + i] += pc[i];
  }
  // the accepted-step count lives in warp 0: publish it to the block
  __syncthreads();
  if (tid == 0) sh->hb[0][0] = (double)R.accepted;
  __syncthreads();
  R.accepted = (long long)sh->hb[0][0];
  R.m = m;
  S.m = m;
  return R;
}

// ------------------------------------------------------------------------------------------------------------
// The same engine spread over W CTAs (a thread-block cluster, or the cooperative grid), for active sets of many
// blocks: with ONE CTA the 32 steps of a block reach the other m - 64 entries through a single SM's L2 port
// (~12 k cycles per block at m = 800, measured), which is what bounds dense active sets.  Here g lives in global
// memory (L2), CTA 0 keeps the chain (warp 0) and the staging (warps 1..15), and EVERY CTA applies the previous
// block's steps to the 32-entry groups it owns (group % W); one barrier of the whole team per block hands over g
// of the next block and h of the finished one.  Same step order per entry as run(): bit-identical iterates.
struct Multi {
  int W, me;      // CTAs in the team, this CTA's index (0 runs the chain)
  double *gG;     // [m] g by entry (global)
  double *hG;     // [2][32] steps of the last two blocks (global)
  double *pmaxG;  // [1]
  int *flagsG;    // [0] new m, [1] list changed
  int *rowG;      // [m] row ids of the list (global; rewritten by CTA 0 when the list is compacted)
};

// apply block hbk's steps to the entries of the 32-entry groups owned by this CTA, skipping blocks ex0 / ex1.
// Warps wfirst.. of the CTA take the owned groups round-robin; lane = entry within the group.
template <int T, class Policy>
__device__ __forceinline__ void apply_owned(const State &S, const Multi &X, int m, int hbk, const double *hsrc, int ex0,
                                            int ex1, int wfirst) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = T / 32 - wfirst;
  if (warp < wfirst) return;
  const int cnt = min(32, m - 32 * hbk);
  const double hl = lane < cnt ? __ldcg(hsrc + lane) : 0.0; // lane i holds h_i
  if (!__any_sync(0xffffffffu, hl != 0.0)) return;
  const int ri = lane < cnt ? gcolidx(S, S.row[S.ord[32 * hbk + lane]]) : 0; // ... and the column of entry i of the block
  const int ngroups = (m + 31) >> 5;
  for (int gi = X.me + X.W * (warp - wfirst); gi < ngroups; gi += X.W * nw) {
    const int t = 32 * gi + lane;
    const bool live = t < m;
    const int blk = live ? (S.pos[t] >> 5) : ex0;
    const bool skip = !live || blk == ex0 || blk == ex1;
    const int rt = live ? S.row[t] : 0;
    double gt = skip ? 0.0 : __ldcg(X.gG + t);
#pragma unroll 1
    for (int i0 = 0; i0 < cnt; i0 += 16) {
      double v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int rr_ = __shfl_sync(0xffffffffu, ri, (i0 + u) & 31);
        v[u] = (!skip && i0 + u < cnt) ? ld_l2(S.G + (long long)rr_ * S.ldg + rt) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const double h = __shfl_sync(0xffffffffu, hl, (i0 + u) & 31);
        if (i0 + u < cnt && h != 0.0) gt = Policy::apply(gt, v[u], h);
      }
    }
    if (!skip) __stcg(X.gG + t, gt);
  }
}

// Team-collective: every thread of every CTA of the team calls it with the same arguments (S.row/ord/pos are
// per-CTA shared-memory copies; S.g is unused, S.be/S.stage/S.sh only matter on CTA 0).  `sync` is the team barrier
// (cluster.sync / grid.sync) with release-acquire semantics on global memory.  Only CTA 0's Result is meaningful.
template <int T, class Policy, class Sync>
__device__ Result run_multi(State &S, const Multi &X, const Policy &P, Sync sync, double rr, long long maxPasses,
                            unsigned long long pass_counter, bool ordered, unsigned long long seed, double optTol,
                            unsigned char *inlist) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool chainCTA = X.me == 0;
  int m = S.m;
  Result R{0, 0, 0, 0.0, 0, m};
  Shared *sh = S.sh;
  // optional cycle counters (S.prof): thread 0 of CTA 0 [0..3] = panel, chain, team barrier, pass ends; thread 32 of
  // CTA 0 [4..6] = staging, apply, team barrier; thread 0 of CTA 1 [7] = apply
  // (compiled in with -DCDGPU_CHAIN_PROF only: the counters cost the naive kernel registers it does not have)
#ifdef CDGPU_CHAIN_PROF
  long long pc[4] = {0, 0, 0, 0};
  const bool prof = S.prof != nullptr && ((chainCTA && (tid == 0 || tid == 32)) || (X.me == 1 && tid == 0));
  long long tp = prof ? clock64() : 0;
  auto lap = [&](int slot) {
    if (prof) {
      const long long t = clock64();
      pc[slot] += t - tp;
      tp = t;
    }
  };
#else
  auto lap = [](int) {};
#endif
  for (long long pass = 0; pass < maxPasses; ++pass) {
    const int m_pass = m;
    const PermKey pkm = cd_perm_key((uint32_t)max(m, 1), seed, pass_counter + pass);
    for (int s = tid; s < m; s += T) {
      const int e = ordered ? s : (int)cd_perm(pkm, (uint32_t)s);
      S.ord[s] = (unsigned short)e;
      S.pos[e] = (unsigned short)s;
    }
    __syncthreads();
    const int nb = (m + 31) >> 5;
    if (chainCTA) {
      stage_block(S, P, m, 0, S.stage, tid, T);
      __syncthreads();
    }
    double pmax = 0.0;
    long long acc = 0;
    lap(3);
    for (int b = 0; b < nb; ++b) {
      if (chainCTA && warp == 0) {
        double *buf = S.stage + (b & 1) * BUF_DOUBLES;
        const int cnt = min(32, m - 32 * b);
        const bool valid = lane < cnt;
        const int e = valid ? S.ord[32 * b + lane] : 0;
        double gj = valid ? __ldcg(X.gG + e) : 0.0, bej = valid ? S.be[e] : 0.0;
        // lanes past the end of the list step on benign constants (their results are never used; garbage could send
        // every step of the warp through a policy's slow path)
        const double c0 = valid ? buf[2048 + lane] : 1.0, c1 = valid ? buf[2048 + 32 + lane] : 1.0, c2 = valid ? buf[2048 + 64 + lane] : 1.0;
        if (b > 0) { // steps of the previous block, in order
          const double *hp = sh->hb[(b - 1) & 1], *Pb = buf + 1024;
#pragma unroll 8
          for (int i = 0; i < 32; ++i) {
            const double h = hp[i];
            if (h != 0.0) gj = Policy::apply(gj, Pb[i * 32 + lane], h);
          }
        }
        lap(0);
        double myh = 0.0;
        double drow = buf[lane];
        for (int i = 0; i < cnt; ++i) {
          const double dnext = buf[((i + 1) & 31) * 32 + lane];
          double nw, hi, dr;
          P.step(gj, bej, c0, c1, c2, rr, nw, hi, dr);
          const double h = __shfl_sync(0xffffffffu, hi, i);
          if (lane == i) {
            bej = nw;
            myh = hi;
          }
          if (h != 0.0) {
            gj = Policy::apply(gj, drow, h);
            if (Policy::HAS_RR) rr += __shfl_sync(0xffffffffu, dr, i);
            acc += 1;
          }
          drow = dnext;
        }
        if (valid) {
          __stcg(X.gG + e, gj);
          S.be[e] = bej;
        }
        const double hv = valid ? myh : 0.0;
        sh->hb[b & 1][lane] = hv;
        __stcg(X.hG + (b & 1) * 32 + lane, hv);
        pmax = fmax(pmax, fabs(myh));
        lap(1);
      } else {
        if (chainCTA && b + 1 < nb) stage_block(S, P, m, b + 1, S.stage + ((b + 1) & 1) * BUF_DOUBLES, tid - 32, T - 32);
        lap(0);
        if (b >= 1) apply_owned<T, Policy>(S, X, m, b - 1, X.hG + ((b - 1) & 1) * 32, b - 1, b, chainCTA ? 1 : 0);
        lap(1);
      }
      sync();
      lap(2);
    }
    // drain: the last block's steps reach the rest of the list (CTA 0's warp 0 joins in)
    if (nb >= 2) apply_owned<T, Policy>(S, X, m, nb - 1, X.hG + ((nb - 1) & 1) * 32, nb - 1, -1, 0);
    // ---- end of the pass on CTA 0: max|h|, dropzeros!
    if (chainCTA) {
      if (warp == 0) {
        pmax = warp_max(pmax);
        if (lane == 0) sh->pmax = pmax;
      }
      int z = 0;
      for (int i = tid; i < m; i += T) z |= (S.be[i] == 0.0);
      z = __syncthreads_or(z);
      R.accepted += __shfl_sync(0xffffffffu, acc, 0);
      if (tid == 0) {
        __stcg(X.pmaxG, sh->pmax);
        __stcg(X.flagsG + 1, z ? 1 : 0);
        if (!z) __stcg(X.flagsG, m);
      }
    }
    sync(); // drain complete everywhere, pass summary published
    const bool changed = __ldcg(X.flagsG + 1) != 0;
    if (changed) { // rare: an entry left the active set — CTA 0 compacts (swap-with-last in list order)
      if (chainCTA) {
        unsigned short *idx = S.ord;
        for (int i = tid; i < m; i += T) idx[i] = (unsigned short)i;
        __syncthreads();
        if (tid == 0) {
          int n = m, i = 0;
          while (i < n) {
            if (S.be[idx[i]] == 0.0) {
              inlist[S.coord[idx[i]]] = 0;
              if (i != n - 1) idx[i] = idx[n - 1];
              n -= 1;
            } else {
              i += 1;
            }
          }
          sh->newm = n;
        }
        __syncthreads();
        const int mn = sh->newm;
        double *tmpd = S.stage;
        int *tmpi = reinterpret_cast<int *>(S.stage);
        for (int i = tid; i < mn; i += T) tmpd[i] = __ldcg(X.gG + idx[i]);
        __syncthreads();
        for (int i = tid; i < mn; i += T) __stcg(X.gG + i, tmpd[i]);
        __syncthreads();
        for (int i = tid; i < mn; i += T) tmpd[i] = S.be[idx[i]];
        __syncthreads();
        for (int i = tid; i < mn; i += T) S.be[i] = tmpd[i];
        __syncthreads();
        for (int i = tid; i < mn; i += T) tmpi[i] = S.row[idx[i]];
        __syncthreads();
        for (int i = tid; i < mn; i += T) {
          S.row[i] = tmpi[i];
          __stcg(X.rowG + i, tmpi[i]);
        }
        __syncthreads();
        if (S.coord != S.row) {
          for (int i = tid; i < mn; i += T) tmpi[i] = S.coord[idx[i]];
          __syncthreads();
          for (int i = tid; i < mn; i += T) S.coord[i] = tmpi[i];
          __syncthreads();
        }
        if (tid == 0) __stcg(X.flagsG, mn);
      }
      sync(); // the compacted list is published
      m = __ldcg(X.flagsG);
      if (!chainCTA) {
        for (int i = tid; i < m; i += T) S.row[i] = __ldcg(X.rowG + i);
        __syncthreads();
      }
    }
    R.npasses += 1;
    R.visits += m_pass;
    const double pm = __ldcg(X.pmaxG);
    R.maxH = pm;
    if (pm < optTol) {
      R.conv = 1;
      break;
    }
  }
  lap(3);
#ifdef CDGPU_CHAIN_PROF
  if (prof) {
    if (X.me == 1) {
      S.prof[7] += pc[1];
    } else {
      const int o = tid == 0 ? 0 : 4;
      for (int i = 0; i < (tid == 0 ? 4 : 3); ++i) S.prof[o + i] += pc[i];
    }
  }
#endif
  if (chainCTA) {
    __syncthreads();
    if (tid == 0) sh->hb[0][0] = (double)R.accepted;
    __syncthreads();
    R.accepted = (long long)sh->hb[0][0];
  }
  R.m = m;
  S.m = m;
  return R;
}

} // namespace chain
