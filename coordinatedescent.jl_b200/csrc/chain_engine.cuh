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
      // every load is issued unconditionally on a clamped (valid) address and masked when it is stored: a predicated
      // load makes the compiler wait for it as soon as the predicate register is needed again, which serialises the batch
      const int idc = min(idx, total - 1);
      const int i = (idc >> 5) & 31, j = idc & 31;
      const bool panel = idc >= 1024;
      const int rj = S.row[S.ord[32 * bb + min(j, cnt - 1)]];
      const int ri = S.row[S.ord[32 * (panel ? bb - 1 : bb) + (panel ? i : min(i, cnt - 1))]];
      v[u] = ld_l2(gcol(S, ri) + rj);
    }
#pragma unroll
    for (int u = 0; u < 5; ++u) {
      const int idx = base + u * nthr;
      if (idx < total) {
        const int i = (idx >> 5) & 31, j = idx & 31;
        buf[idx] = (j < cnt && (idx >= 1024 || i < cnt)) ? v[u] : 0.0;
      }
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
        v[u] = ld_l2(gcol(S, S.row[S.ord[32 * hbk + min(i0 + u, cnt - 1)]]) + rt); // unconditional (see stage_block); h = 0 past the end
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const double h = (i0 + u < cnt) ? hv[i0 + u] : 0.0;
        if (h != 0.0) gt = Policy::apply(gt, v[u], h);
      }
    }
    S.g[t] = gt;
  }
}

// The 32 dependent steps of one block on warp 0 (lane j owns entry j of the block; D = the staged 32x32 diagonal block,
// D[i*32 + j] = G(entry j, entry i)).  Policies with FAST_V (soft-threshold steps: LS / WLS / quadratic) carry, next to
// the gradient g_j, the argument of the shrinkage v_j = be_j -+ g_j/a_jj itself: it is formed once per block exactly as
// step() forms it, and a step h on entry i moves it by one fused multiply-add with the pre-scaled row D[j,i]/a_jj, so
// that the dependent path of a step is FMA -> compare/select -> subtract -> shuffle (~60 cycles) instead of
// multiply, add, quotient (3 FMAs + range test), add, shrink, subtract, shuffle (~150).  g_j receives the same
// unfused updates as before, off the dependent path; v_j is re-derived from it at the next block, so the two never
// drift apart by more than the 32 roundings of a block.
template <class Policy>
__device__ __forceinline__ void chain_steps(const Policy &P, const double *buf, int cnt, int lane, double &gj, double &bej,
                                            double c0, double c1, double c2, double &rr, double &myh, long long &acc) {
  double drow = buf[lane];
  if constexpr (Policy::FAST_V) {
    double v = P.enter(gj, bej, c0, c1);
    double ds = lane > 0 || cnt != 32 ? __dmul_rn(drow, c1) : 0.0; // (full block: lane 0 is frozen from the start)
    // nw = shrink(v, c2), h = nw - be with both branches evaluated ahead of the comparisons
    auto prox = [&](double vv, double &nw, double &hi) {
      const double up = __dsub_rn(vv, c2), dn = __dadd_rn(vv, c2); // (intrinsics: kept as two adds next to the compares)
      const double hup = __dsub_rn(up, bej), hdn = __dsub_rn(dn, bej);
      const bool gt = vv > c2, lt = vv < -c2;
      nw = gt ? up : (lt ? dn : 0.0);
      hi = gt ? hup : (lt ? hdn : -bej);
    };
    if (cnt == 32) {
      // full block, straight-line: constant shared-memory offsets and shuffle lanes; lane i's v is frozen from its own
      // step on (lanes <= i multiply the step by a zero row), so that its step is read off v once more after the loop
      // instead of being captured by four selects in every step
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const double dnext = buf[((i + 1) & 31) * 32 + lane];
        double nw, hi;
        prox(v, nw, hi);
        const double h = __shfl_sync(0xffffffffu, hi, i);
        v = fma(-ds, h, v);
        gj = Policy::apply(gj, drow, h); // h == 0: the product is +-0 and g keeps its value
        drow = dnext;
        ds = lane > i + 1 ? __dmul_rn(dnext, c1) : 0.0; // a zero row freezes v (the select is off the dependent path)
      }
      double nw;
      prox(v, nw, myh);
      bej = nw;
      acc += __popc(__ballot_sync(0xffffffffu, myh != 0.0));
    } else {
      int nacc = 0;
      for (int i = 0; i < cnt; ++i) {
        const double dnext = buf[((i + 1) & 31) * 32 + lane];
        double nw, hi;
        prox(v, nw, hi);
        const double h = __shfl_sync(0xffffffffu, hi, i);
        v = fma(-ds, h, v); // (lane i itself: its v is not read again in this block)
        if (lane == i) {
          bej = nw;
          myh = hi;
        }
        gj = Policy::apply(gj, drow, h);
        nacc += h != 0.0;
        drow = dnext;
        ds = __dmul_rn(dnext, c1);
      }
      acc += nacc;
    }
  } else {
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
        chain_steps(P, buf, CHAIN_PROBE_BIT(2) ? 0 : cnt, lane, gj, bej, c0, c1, c2, rr, myh, acc);
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
// The same engine spread over W >= 2 CTAs (a thread-block cluster, or the cooperative grid), for active sets of many
// blocks: with ONE CTA the 32 steps of a block reach the other m - 64 entries through a single SM's L2 port
// (~12 k cycles per block at m = 800, measured), which is what bounds dense active sets.  CTA 0 keeps the chain
// (warp 0) and the staging (warps 1..15) and nothing else.  The 32-entry groups of the list are dealt to the warps of
// the other CTAs (group gi: CTA 1 + gi % (W-1), warp gi / (W-1)); an owner applies the steps of block b to its groups as
// soon as they are published.  Inside a pass there is NO barrier and NO fence between the CTAs: g and h travel through
// global memory (L2) as TAGGED values — 16 bytes {low word, tag, high word, tag}, each 8-byte half written and read
// atomically, the tag = number of blocks (counted over all passes of the call) whose steps the value contains — and
// every reader knows the tag it needs, so it simply re-reads until both halves carry it:
//   h of block b            tag base+b+1, written by the chain warp, awaited by every owner;
//   g of an entry, owners   an owner applying block b reads tag base+b and writes base+b+1; it leaves the entries of
//                           blocks b and b+1 alone (block b+1 gets the steps through the staged panel);
//   g of an entry, chain    before block b the chain warp needs tag >= base+b-1 (the owners have the whole chain of
//                           block b-1 to get there: their L2 round trips are off the dependent path) and writes base+b+1.
// One team barrier per pass (max|h| and the list state published).  Every entry still receives the steps in visit
// order: same iterates as run().
struct Multi {
  int W, me;      // CTAs in the team, this CTA's index (0 runs the chain)
  double *gG;     // [m] g by entry (global): input and output of the call
  double *hG;     // [64] unused by the tagged pipeline (kept: callers lay pmaxG / flagsG out behind it)
  double *pmaxG;  // [1]
  int *flagsG;    // [0] new m, [1] list changed
  int *rowG;      // [m] row ids of the list (global; rewritten by CTA 0 when the list is compacted)
  uint4 *gT;      // [m] tagged g by entry (global)
  uint4 *hT;      // [ceil32(m)] tagged steps of the current pass by visit position (global)
};

__device__ __forceinline__ void tag_store(uint4 *p, double x, unsigned tag) {
  const unsigned lo = (unsigned)__double2loint(x), hi = (unsigned)__double2hiint(x);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(lo), "r"(tag), "r"(hi), "r"(tag) : "memory");
}
// re-read until both halves carry the same tag >= need
__device__ __forceinline__ double tag_wait(const uint4 *p, unsigned need) {
  unsigned lo, t0, hi, t1;
  do {
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(t0), "=r"(hi), "=r"(t1) : "l"(p) : "memory");
  } while (t0 != t1 || (int)(t0 - need) < 0);
  return __hiloint2double((int)hi, (int)lo);
}

// One pass of an owner warp (every warp of the CTAs 1..W-1): its groups are (me-1) + (W-1)*warp, + (W-1)*NW, ...;
// block by block as the steps are published.  A warp with ONE group (the usual case) that has a 32x32 tile of the CTA's
// idle stage area (warps 0..3) fetches its G entries for the next block into that tile while the chain of that block
// runs, so that between "h is published" and "g is handed on" there is one L2 read and 32 dependent updates.
#ifdef CDGPU_CHAIN_PROF
#define OWNER_LAP(slot)                                    \
  do {                                                     \
    if (prof_apply && tid == 0) {                          \
      const long long t_ = clock64();                      \
      if ((slot) >= 0) prof_apply[(slot)] += t_ - tlap;    \
      tlap = t_;                                           \
    }                                                      \
  } while (0)
#else
#define OWNER_LAP(slot) \
  do {                  \
  } while (0)
#endif
template <int T, class Policy>
__device__ __forceinline__ void owner_pass(const State &S, const Multi &X, int m, int nb, unsigned base, long long *prof_apply) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = T / 32;
  const int g0 = (X.me - 1) + (X.W - 1) * warp, gstep = (X.W - 1) * NW;
  if (g0 >= nb) return;
  long long tacc = 0;
#ifdef CDGPU_CHAIN_PROF
  long long tlap = 0; // prof_apply[1..5]: wait for g, wait for h, 32 updates, tagged store, prefetch of the next block
#endif
  const bool tiled = g0 + gstep >= nb && warp < STAGE_DOUBLES / 1024;
  double *tile = S.stage + 1024 * warp; // [i*32 + lane] = G(entry t1, entry i of the block)
  const int t1 = 32 * g0 + lane;
  const bool live1 = t1 < m;
  const int blk1 = live1 ? (S.pos[t1] >> 5) : -1, rt1 = live1 ? S.row[t1] : 0;
  auto prefetch = [&](int b) {
    const int cnt = min(32, m - 32 * b);
    const int ri = lane < cnt ? gcolidx(S, S.row[S.ord[32 * b + lane]]) : 0;
    const bool skip = !live1 || blk1 == b || blk1 == b + 1;
#pragma unroll 1
    for (int i0 = 0; i0 < 32; i0 += 16) {
      double v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) { // unconditional loads on valid addresses (ri = 0 past the block, rt1 = 0 past the list)
        const int rr_ = __shfl_sync(0xffffffffu, ri, i0 + u);
        v[u] = ld_l2(S.G + (long long)rr_ * S.ldg + rt1);
      }
#pragma unroll
      for (int u = 0; u < 16; ++u) tile[(i0 + u) * 32 + lane] = (!skip && i0 + u < cnt) ? v[u] : 0.0;
    }
  };
  if (tiled) prefetch(0);
  for (int b = 0; b < nb; ++b) {
    const unsigned tagb = base + (unsigned)b; // what the entries carry before this block; + 1 afterwards
    if (tiled) {
      const bool skip = !live1 || blk1 == b || blk1 == b + 1;
      OWNER_LAP(-1);
      double gt = skip ? 0.0 : tag_wait(X.gT + t1, tagb);          // own last write, or the chain warp's (block b-1)
      OWNER_LAP(1);
      const double hl = tag_wait(X.hT + 32 * b + lane, tagb + 1u); // lane i holds h_i
      __syncwarp();
      OWNER_LAP(2);
      const long long tp0 = prof_apply ? clock64() : 0;
#pragma unroll 8
      for (int u = 0; u < 32; ++u) gt = Policy::apply(gt, tile[u * 32 + lane], __shfl_sync(0xffffffffu, hl, u)); // (skip: tile = 0)
      OWNER_LAP(3);
      if (!skip) tag_store(X.gT + t1, gt, tagb + 1u);
      OWNER_LAP(4);
      if (b + 1 < nb) prefetch(b + 1);
      OWNER_LAP(5);
      if (prof_apply) tacc += clock64() - tp0;
    } else {
      const double hl = tag_wait(X.hT + 32 * b + lane, tagb + 1u);
      const long long tp0 = prof_apply ? clock64() : 0;
      const int cnt = min(32, m - 32 * b);
      const int ri = lane < cnt ? gcolidx(S, S.row[S.ord[32 * b + lane]]) : 0;
      for (int gi = g0; gi < nb; gi += gstep) {
        const int t = 32 * gi + lane;
        const bool live = t < m;
        const int blk = live ? (S.pos[t] >> 5) : b;
        const bool skip = !live || blk == b || blk == b + 1;
        const int rt = live ? S.row[t] : 0;
        double gt = skip ? 0.0 : tag_wait(X.gT + t, tagb);
        __syncwarp();
#pragma unroll 1
        for (int i0 = 0; i0 < cnt; i0 += 16) {
          double v[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const int rr_ = __shfl_sync(0xffffffffu, ri, (i0 + u) & 31);
            v[u] = ld_l2(S.G + (long long)rr_ * S.ldg + rt); // unconditional, valid address (see stage_block)
          }
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const double h = __shfl_sync(0xffffffffu, hl, (i0 + u) & 31); // (0 past the end of the block)
            gt = Policy::apply(gt, v[u], h);
          }
        }
        if (!skip) tag_store(X.gT + t, gt, tagb + 1u);
      }
      if (prof_apply) tacc += clock64() - tp0;
    }
  }
  if (prof_apply && tid == 0) *prof_apply += tacc;
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
  // optional cycle counters (S.prof): thread 0 of CTA 0 [0..3] = panel, chain, CTA barrier, pass ends, [5] = wait for the
  // owners' flags; thread 32 of CTA 0 [4], [6] = staging, CTA barrier; thread 0 of CTA 1 [7] = apply
  // (compiled in with -DCDGPU_CHAIN_PROF only: the counters cost the naive kernel registers it does not have)
#ifdef CDGPU_CHAIN_PROF
  long long pc[5] = {0, 0, 0, 0, 0};
  const bool prof = S.prof != nullptr && ((chainCTA && (tid == 0 || tid == 32)) || (X.me == 1 && tid == 0));
  long long tp = prof ? clock64() : 0;
  auto lap = [&](int slot) {
    if (prof) {
      const long long t = clock64();
      pc[slot] += t - tp;
      tp = t;
    }
  };
  long long *prof_apply = (S.prof != nullptr && X.me == 1) ? S.prof + 7 : nullptr; // thread 0 of CTA 1: cycles applying
#else
  auto lap = [](int) {};
  long long *prof_apply = nullptr;
#endif
  // tagged copies of g (tag 0); one team barrier per call
  for (int i = X.me * T + tid; i < ((m + 31) & ~31); i += X.W * T) {
    if (i < m) tag_store(X.gT + i, __ldcg(X.gG + i), 0u);
    tag_store(X.hT + i, 0.0, 0u); // (tags left by an earlier call must not pass for this call's)
  }
  __threadfence();
  sync();
  unsigned base = 0; // blocks of the passes before this one
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
    double pmax = 0.0;
    long long acc = 0;
    if (chainCTA) {
      stage_block(S, P, m, 0, S.stage, tid, T);
      __syncthreads();
      lap(3);
      for (int b = 0; b < nb; ++b) {
        if (warp == 0) {
          double *buf = S.stage + (b & 1) * BUF_DOUBLES;
          const int cnt = min(32, m - 32 * b);
          const bool valid = lane < cnt;
          const int e = valid ? S.ord[32 * b + lane] : 0;
          // g with the steps of every block <= b-2 of this pass (and of all earlier passes) applied
          double gj = valid ? tag_wait(X.gT + e, base + (unsigned)max(b - 1, 0)) : 0.0, bej = valid ? S.be[e] : 0.0;
          __syncwarp();
          lap(4);
          // lanes past the end of the list step on benign constants (their results are never used; garbage could send
          // every step of the warp through a policy's slow path)
          const double c0 = valid ? buf[2048 + lane] : 1.0, c1 = valid ? buf[2048 + 32 + lane] : 1.0, c2 = valid ? buf[2048 + 64 + lane] : 1.0;
          if (b > 0) { // steps of the previous block, in order
            const double *hp = sh->hb[(b - 1) & 1], *Pb = buf + 1024;
            // operands first, then the 32 dependent updates (a zero step leaves gj as it is: the skipped product is +-0)
#pragma unroll
            for (int i0 = 0; i0 < 32; i0 += 16) {
              double pv[16], hv[16];
#pragma unroll
              for (int u = 0; u < 16; ++u) {
                pv[u] = Pb[(i0 + u) * 32 + lane];
                hv[u] = hp[i0 + u];
              }
#pragma unroll
              for (int u = 0; u < 16; ++u) gj = Policy::apply(gj, pv[u], hv[u]);
            }
          }
          lap(0);
          double myh = 0.0;
          chain_steps(P, buf, cnt, lane, gj, bej, c0, c1, c2, rr, myh, acc);
          const double hv = valid ? myh : 0.0;
          tag_store(X.hT + 32 * b + lane, hv, base + (unsigned)b + 1u); // all 32 lanes: owners wait on every lane's tag
          if (valid) {
            tag_store(X.gT + e, gj, base + (unsigned)b + 1u);
            S.be[e] = bej;
          }
          sh->hb[b & 1][lane] = hv;
          pmax = fmax(pmax, fabs(myh));
          lap(1);
        } else {
          if (b + 1 < nb) stage_block(S, P, m, b + 1, S.stage + ((b + 1) & 1) * BUF_DOUBLES, tid - 32, T - 32);
          lap(0);
        }
        __syncthreads(); // the next block is staged, hb[b & 1] is visible to the next panel
        lap(2);
      }
    } else {
      owner_pass<T, Policy>(S, X, m, nb, base, prof_apply);
    }
    base += (unsigned)nb;
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
        for (int i = tid; i < mn; i += T) tmpd[i] = tag_wait(X.gT + idx[i], base); // (every entry ends a pass with tag base)
        __syncthreads();
        for (int i = tid; i < mn; i += T) tag_store(X.gT + i, tmpd[i], base);
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
    } else {
      const int o = tid == 0 ? 0 : 4;
      for (int i = 0; i < (tid == 0 ? 4 : 3); ++i)
        if (tid == 0 || i != 1) S.prof[o + i] += pc[i];
      if (tid == 0) S.prof[5] += pc[4]; // the chain warp's wait for the owners' flags
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
