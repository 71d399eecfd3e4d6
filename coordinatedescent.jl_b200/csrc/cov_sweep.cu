// cov_sweep.cu — covariance-form (CDQuadraticLoss) active-set coordinate descent, whole
// lambda path in ONE launch of ONE thread-block cluster.
//
// Reference semantics: _coordinateDescent!/_cdPass! (src/coordinate_descent.jl:65-110) driving
// descendCoordinate!(::CDQuadraticLoss) (src/cd_differentiable_function.jl:324-348):
//     x_k <- S(x_k - (Ax_k + b_k)/A_kk, lambda0*omega_k/A_kk);   Ax += A[:,k] * h.
//
// B200 design:
//  * the p coordinates are partitioned over the C CTAs of a cluster (C = 16, non-portable size);
//    each CTA keeps its slice of Ax, b, 1/diag(A), omega, beta in shared memory for the whole path.
//  * FULL pass = exact Gauss-Seidel by speculation on WHO moves.  In a full pass the coordinates that move
//    are, almost always, the current members of the iterate; everybody else is a no-op whose only role is
//    "am I still inactive when my turn comes?".  So a pass is (1) the member chain: CTA 0 runs the members, in
//    visit order, through the blocked chain engine (chain_engine.cuh) assuming no other coordinate moves — a
//    dependent chain over the |members| x |members| block of A only; (2) verification, all CTAs, bandwidth
//    bound: every coordinate j accumulates Ax_j += A[j,k]*h_k over the members in visit order (the same
//    non-fused multiply-adds, in the same order, as the one-step-at-a-time reference) and every non-member
//    is tested at the point of the accumulation where its own visit falls; (3) one candidate exchange over
//    DSMEM (st.async + mbarrier) gives the FIRST non-member that moves, if any: the chain is valid up to that
//    position, the entering step is applied, and the pass resumes behind it with the remaining members.
//    Without an entering coordinate (the common case) the whole pass costs one chain and one sweep over
//    |members| columns; iterates, visit counts and the SparseIterate order are those of the reference.
//  * ACTIVE-SET passes only ever read Ax on the active set, so they run inside CTA 0 on the chain engine
//    (or, for many 32-entry blocks, distributed over the cluster); the other CTAs sleep on the cluster
//    barrier.  Afterwards the whole cluster folds the accumulated change into its Ax slices
//    (Ax += A[:, act] * (beta - beta_at_entry), a coalesced GEMV over L2-resident columns).
//  * lists longer than the engine holds fall back to an event-by-event pass (first mover by cluster-wide
//    candidate exchange, one column axpy per step): slow, exact, no size limit.
//  * columns of A may be formed lazily (colslot/resume in CovArgs): the kernel leaves at an entering coordinate
//    whose column is missing and continues, in a later launch, exactly where it stopped.
//  * no grid-wide sync, no host round trip between passes or between lambdas.
#include <cooperative_groups.h>

#include <algorithm>

#include "chain_engine.cuh"
#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int COV_T = 512;
constexpr int COV_ACT_CAP = 4096; // most entries the in-CTA engine is ever given (launcher may pick less)
constexpr int COV_MAXC = 16;
constexpr unsigned KEY_NONE = 0xffffffffu;

struct __align__(16) Cand { // 32 bytes = two st.async.v2.b64
  unsigned key; // visit position of the coordinate (KEY_NONE: this slice has no mover)
  int k;        // coordinate | member flag in bit 31
  int pad;      // event pass: sender's count of visited-but-not-appended non-members; chain pass: "some test saw v == 0"
  int unused;
  double h, nw;
};

struct Bcast { // CTA 0 -> cluster after an active phase / at the end of a lambda
  long long npasses, visits, accepted;
  double maxH;
  int m0, nact, status, conv;
};

struct Smem {
  Cand cand[2][COV_MAXC];
  Cand wcand[COV_T / 32]; // per-warp best of the current scan
  Bcast bc;
  double h[2];
  int nact, flag, nonapp;
  int mR, tz;             // chain pass: entries handed to the member chain this round; sticky "a test saw v == 0"
  int nd, nc;             // screening: entries of D (CTA 0); CURRENT rows of this CTA
  double Dn2;             // screening: squared distance bound of this round (CTA 0)
  int s2[2];
  unsigned long long mbar[2]; // candidate-exchange barriers (one per round parity)
  chain::Shared ch;
};

__device__ __forceinline__ void named_bar(int id, int nthr) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthr) : "memory"); }

struct Ctx {
  const CovArgs &a;
  cg::cluster_group &cluster;
  Smem *sm;
  double *sAx, *sAx2, *sb, *sainv, *sw, *sbeta; // this CTA's slice (shared or global); sAx2: tentative Ax of a chain pass
  int *s_act;                            // CTA-0 engine: coordinates of the stored entries
  double *e_g, *e_be, *e_h, *e_stage;    // CTA-0 engine state (chain_engine.cuh); e_h: steps of the member chain
  unsigned *e_vpos;                      // visit positions of the member chain's entries
  unsigned short *e_ord, *e_pos;
  int ecap;                              // entries the engine can hold in this launch
  int multi_ok;                          // the cluster-distributed engine may be used (CDGPU_COV_MULTI=0 disables)
  unsigned char *s_in, *s_vnz;           // per slice: member at pass start / tentative value non-zero
  int rank, C, L, lo, len, slice_in_smem;
};

// start of column k of A (lazily formed columns live in slots)
__device__ __forceinline__ const double *col_ptr(const CovArgs &a, int k) {
  return a.A + (long long)(a.colslot ? __ldg(a.colslot + k) : k) * a.lda;
}

// remote (or own) element of a slice array living at the same shared offset in every CTA
__device__ __forceinline__ double slice_get(const Ctx &c, double *arr_local, int k) {
  int owner = k / c.L;
  if (c.slice_in_smem) {
    const double *r = c.cluster.map_shared_rank(arr_local, owner);
    return r[k - owner * c.L];
  }
  // global slices: arr_local == base + lo
  return __ldcg(arr_local - c.lo + k);
}

// ---------------------------------------------------------- DSMEM message passing --
// The per-step candidate exchange of a full pass does not go through a cluster barrier (whose
// release semantics cost a GPU-scope MEMBAR per round): every CTA pushes its 32-byte candidate into
// each peer's shared memory with st.async, which also signals the peer's mbarrier (complete_tx);
// a CTA only waits on its OWN mbarrier until all C candidates of the round have landed.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(addr), "r"(parity)
                 : "memory");
  } while (!done);
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async_16(uint32_t raddr, unsigned long long a, unsigned long long b, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];" ::"r"(raddr),
               "l"(a), "l"(b), "r"(rbar)
               : "memory");
}

// One round of the cluster-wide "first mover" election.  Every thread brings its best candidate (best == KEY_NONE:
// none); on return every thread of every CTA holds the cluster's winner (key == KEY_NONE: nobody moves) and the sum
// of the CTAs' `pad` words.  One __syncthreads, one st.async push per peer, one wait on the CTA's own mbarrier.
__device__ __forceinline__ Cand elect(Ctx &c, unsigned &round, unsigned best, int bk, double bh, double bnw, int pad,
                                      int &pad_sum) {
  Smem *sm = c.sm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, C = c.C;
  const unsigned par = round & 1u, phase = (round >> 1) & 1u;
  if (tid == 0) mbar_arrive_expect_tx(&sm->mbar[par], (uint32_t)(C * sizeof(Cand)));
  const unsigned wmin = __reduce_min_sync(0xffffffffu, best);
  if (best == wmin && (best != KEY_NONE || lane == 0)) { // positions are unique: one lane per warp
    Cand &wc = sm->wcand[warp];
    wc.key = best;
    wc.k = bk;
    wc.h = bh;
    wc.nw = bnw;
  }
  __syncthreads();
  if (warp == 0) {
    const unsigned mykey = lane < COV_T / 32 ? sm->wcand[lane].key : KEY_NONE;
    const unsigned bmin = __reduce_min_sync(0xffffffffu, mykey);
    const unsigned src = __ffs(__ballot_sync(0xffffffffu, mykey == bmin && lane < COV_T / 32)) - 1; // winning warp
    if (lane < C) { // push the block's candidate into peer `lane`'s slot [par][rank] and signal its mbarrier
      const Cand &bc = sm->wcand[src];
      const unsigned long long w0 = (unsigned long long)bmin | ((unsigned long long)(unsigned)bc.k << 32);
      const unsigned long long w1 = (unsigned long long)(unsigned)pad;
      const uint32_t slot = map_to_rank(smem_u32(&sm->cand[par][c.rank]), (uint32_t)lane);
      const uint32_t rbar = map_to_rank(smem_u32(&sm->mbar[par]), (uint32_t)lane);
      st_async_16(slot, w0, w1, rbar);
      st_async_16(slot + 16, (unsigned long long)__double_as_longlong(bc.h), (unsigned long long)__double_as_longlong(bc.nw), rbar);
    }
  }
  mbar_wait(&sm->mbar[par], phase);
  // every warp picks the cluster's winner lane-parallel: lane q looks at CTA q's candidate
  const unsigned ckey = lane < C ? sm->cand[par][lane].key : KEY_NONE;
  int napp = lane < C ? sm->cand[par][lane].pad : 0;
  const unsigned wkey = __reduce_min_sync(0xffffffffu, ckey);
  pad_sum = __reduce_add_sync(0xffffffffu, napp);
  const unsigned wsrc = __ffs(__ballot_sync(0xffffffffu, ckey == wkey && lane < C)) - 1;
  Cand w = sm->cand[par][wsrc];
  w.key = wkey;
  round += 1;
  return w;
}

// ------------------------------------------------------------------ event-by-event pass --
// The pass as a sequence of cluster-wide "who moves first?" elections, one column axpy per step.  Used for full passes
// when the list does not fit the chain engine (and as a diagnostic, CDGPU_COV_EVENTS=1), and — RESTRICT — for
// active-set passes over lists longer than the engine holds: then only list members are visited, in list order
// (lpos[k] = list position of coordinate k; random iterator: the keyed permutation of the m list positions).
// Returns max|h|.  Full pass: on return CTA 0 holds the new entries appended (in visit order) behind the m_old old
// ones in a.act; `nonapp_total` is the number of visited non-members whose tentative value was exactly zero (they are
// NOT appended by the reference's setindex!; practically always 0).
template <bool PROF, bool RESTRICT>
__device__ __forceinline__ double event_pass(Ctx &c, double lam, unsigned long long pass_counter, unsigned &round,
                                             long long &accepted, bool first_pass_of_kernel, int &nonapp_total,
                                             long long *pf, int &m_bound, const int *lpos, int m_list) {
  const CovArgs &a = c.a;
  Smem *sm = c.sm;
  const int tid = threadIdx.x;
  const bool ordered = a.randomize == 0;
  const PermKey pk = cd_perm_key((uint32_t)(RESTRICT ? max(m_list, 1) : a.p), a.seed, pass_counter);
  // hot-loop state in registers
  double *__restrict__ sAx = c.sAx;
  const double *__restrict__ sb = c.sb, *__restrict__ sainv = c.sainv, *__restrict__ sw = c.sw;
  double *__restrict__ sbeta = c.sbeta;
  unsigned char *__restrict__ s_in = c.s_in, *__restrict__ s_vnz = c.s_vnz;
  const int lo = c.lo, len = c.len, rank = c.rank;
  int cur = -1;          // ordered full pass: last coordinate that moved
  long long curpos = -1; // otherwise: its visit position
  double maxH = 0.0;
  // membership at the start of the pass (beta != 0, or an explicit zero handed in by the caller)
  for (int i = tid; i < len; i += COV_T) {
    s_in[i] = first_pass_of_kernel ? a.inlist[lo + i] : (unsigned char)(sbeta[i] != 0.0);
    s_vnz[i] = 1;
  }
  if (tid == 0) sm->nonapp = 0;
  __syncthreads();
  for (;;) {
    const long long ta = PROF ? clock64() : 0;
    unsigned best = KEY_NONE;
    int bk = 0;
    double bh = 0.0, bnw = 0.0;
    const int i0 = (ordered && !RESTRICT) ? max(0, cur + 1 - lo) : 0;
    // every thread owns the slice elements tid, tid+T, ... in BOTH the scan and the update below, so no
    // barrier is needed between a step and the next scan
#pragma unroll 2
    for (int i = tid; i < len; i += COV_T) {
      if (i < i0) continue;
      const bool member = s_in[i] != 0;
      unsigned key;
      if (RESTRICT) {
        if (!member) continue;
        const unsigned lp = (unsigned)__ldcg(lpos + lo + i);
        key = ordered ? lp : cd_perm_inv(pk, lp);
        if ((long long)key <= curpos) continue;
      } else {
        key = ordered ? (unsigned)(lo + i) : cd_perm_inv(pk, (uint32_t)(lo + i));
        if (!ordered && (long long)key <= curpos) continue;
      }
      const double ainv = sainv[i];
      const double g = sAx[i] + sb[i];
      const double old = sbeta[i];
      const double t = __dmul_rn(g, ainv);
      const double thr = __dmul_rn(__dmul_rn(ainv, lam), sw[i]);
      // short dependent chain for the common case (x_k == 0 stays 0): v = -t, and S(v, thr) != 0 <=> |t| > thr;
      // t == 0 is the (practically impossible) "not appended" case tracked below
      if (old == 0.0 && !(fabs(t) > thr) && t != 0.0 && s_vnz[i]) continue;
      const double v = __dsub_rn(old, t);
      const double nw = cd_shrink(v, thr);
      const double h = nw - old;
      const unsigned char nz = (unsigned char)(v != 0.0); // `x[k] -= b*a` appends iff the value is non-zero
      if (nz != s_vnz[i]) {
        s_vnz[i] = nz;
        if (!member) atomicAdd(&sm->nonapp, nz ? -1 : 1);
      }
      if (h != 0.0 && key < best) {
        best = key;
        bk = (lo + i) | (member ? (int)0x80000000 : 0);
        bh = h;
        bnw = nw;
      }
    }
    const long long tb = PROF ? clock64() : 0;
    int napp = 0;
    // sm->nonapp is read by warp 0 after elect()'s __syncthreads; all atomics above precede it
    __syncthreads();
    const Cand w = elect(c, round, best, bk, bh, bnw, sm->nonapp, napp);
    const long long tc = PROF ? clock64() : 0;
    if (PROF) pf[7] += tb - ta;
    if (PROF) pf[8] += tc - tb;
    if (w.key == KEY_NONE) {
      nonapp_total = napp; // every slice's scan was final
      break;
    }
    const int k = w.k & 0x7fffffff;
    const bool wmember = w.k < 0;
    // column slice first (independent loads in flight), then the shared-memory update
    const double *col = col_ptr(a, k) + lo;
    for (int i = tid; i < len; i += 4 * COV_T) {
      double x[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) x[u] = (i + u * COV_T < len) ? __ldg(col + i + u * COV_T) : 0.0;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i + u * COV_T < len) sAx[i + u * COV_T] = __dadd_rn(sAx[i + u * COV_T], __dmul_rn(x[u], w.h));
    }
    if (k >= lo && k < lo + len && ((k - lo) % COV_T) == tid) { // the thread that scans this element
      sbeta[k - lo] = w.nw;
      s_in[k - lo] = 1;
    }
    if (rank == 0 && tid == 0 && !wmember) { // setindex! appends on the first non-zero store
      a.act[sm->nact] = k;
      sm->nact += 1;
    }
    if (!wmember) m_bound += 1; // every CTA: upper bound of the list length (dropzeros! can only shorten it)
    maxH = fmax(maxH, fabs(w.h));
    accepted += 1;
    cur = k;
    curpos = (long long)w.key;
    if (PROF) pf[9] += clock64() - tc;
  }
  return maxH;
}


// ------------------------------------------------------ active-set engine (CTA 0) --
// Consecutive active-set passes until one has maxH < optTol or `maxPasses` are used: the blocked
// warp-level chain of chain_engine.cuh on (Ax_t, beta_t) of the stored entries, G = A[act, act]
// gathered from the L2-resident columns of A.
struct CovPolicy {
  static constexpr bool HAS_RR = false;
  static constexpr bool FAST_V = true; // chain_engine.cuh: chain_steps
  // argument of the shrinkage, as step() forms it
  __device__ __forceinline__ double enter(double g, double be, double c0, double c1) const {
    return __dsub_rn(be, __dmul_rn(g + c0, c1));
  }
  const double *b, *ainv, *omega;
  double lam;
  __device__ __forceinline__ void load_consts(int k, double &c0, double &c1, double &c2) const {
    c0 = __ldg(b + k);
    c1 = __ldg(ainv + k);
    const double w = omega ? __ldg(omega + k) : 1.0;
    c2 = __dmul_rn(__dmul_rn(c1, lam), w);
  }
  // cd_differentiable_function.jl:330-337
  __device__ __forceinline__ void step(double g, double be, double c0, double c1, double c2, double, double &nw, double &h,
                                       double &dr) const {
    const double gg = g + c0;
    const double v = __dsub_rn(be, __dmul_rn(gg, c1));
    nw = cd_shrink(v, c2);
    h = nw - be;
    dr = 0.0;
  }
  static __device__ __forceinline__ double apply(double g, double Gv, double h) { return __dadd_rn(g, __dmul_rn(Gv, h)); } // :343-345
};

// ------------------------------------------------------------------ full pass: member chain + verification --
// visit position of coordinate k in the full pass
__device__ __forceinline__ unsigned visit_pos(bool ordered, const PermKey &pk, int k) {
  return ordered ? (unsigned)k : cd_perm_inv(pk, (uint32_t)k);
}

// Ax2[slice] = Ax[slice] + sum_{t < tend} A[slice, k_t] * h_t (steps in order, non-fused) [+ A[slice, kx] * hx].
// TEST: every still-unvisited non-member is tested where its own visit falls among the steps; returns this
// thread's earliest mover and records "value exactly zero" in s_vnz.  Inlined so that the compiler keeps the address
// spaces (shared-memory arrays as LDS, not generic loads competing with the column loads in the LSU queue).
struct SweepIn {
  const double *A;
  long long lda;
  const int *colslot;
  const double *sAx, *sb, *sainv, *sw, *sbeta;
  double *sAx2;
  const int *rk;
  const double *rh;
  const unsigned *rpos;
  unsigned char *s_in, *s_vnz;
  int *tz;
  int lo, len, tend, kx, ordered;
  double lam, hx;
  long long curpos;
  PermKey pk;
  long long *pfp; // optional profile slots [4]: set-up, sweep loop, tests, tail
};
struct SweepOut {
  unsigned best;
  int bk;
  double bh, bnw;
};
template <bool TEST>
__device__ __forceinline__ SweepOut sweep_members(const SweepIn &c) {
  const int tid = threadIdx.x, lo = c.lo, len = c.len, tend = c.tend;
  const double *__restrict__ sAx = c.sAx;
  double *__restrict__ sAx2 = c.sAx2;
  const int *__restrict__ rk = c.rk;
  const double *__restrict__ rh = c.rh;
  const unsigned *__restrict__ rpos = c.rpos;
  SweepOut o{KEY_NONE, 0, 0.0, 0.0};
  constexpr int NE = 3, NQ = 4; // elements per thread and chain steps per batch: NE*NQ independent loads in flight
  int tzseen = 0;
  for (int base = 0; base < len; base += NE * COV_T) {
    const long long q0 = c.pfp ? clock64() : 0;
    double acc[NE];
    int ti[NE]; // step index before which the element is tested (-1: never)
    unsigned pj[NE];
#pragma unroll
    for (int u = 0; u < NE; ++u) {
      const int i = base + tid + u * COV_T;
      const bool valid = i < len;
      acc[u] = valid ? sAx[i] : 0.0;
      ti[u] = -1;
      pj[u] = 0;
      if (TEST && valid && !c.s_in[i]) {
        pj[u] = visit_pos(c.ordered != 0, c.pk, lo + i);
        if ((long long)pj[u] > c.curpos) { // lower bound: number of chain entries visited before this coordinate
          int l = 0, r = tend;
          while (l < r) {
            const int mid = (l + r) >> 1;
            if (rpos[mid] < pj[u]) l = mid + 1; else r = mid;
          }
          ti[u] = l;
        }
      }
    }
    double accv[NE]; // the running sum at the moment of the element's own visit (captured by a select: no divergence)
#pragma unroll
    for (int u = 0; u < NE; ++u) accv[u] = acc[u];
    auto test = [&](int u) {
      const int i = base + tid + u * COV_T;
      const double ainv = c.sainv[i];
      const double g = accv[u] + c.sb[i];
      const double old = c.sbeta[i];
      const double t = __dmul_rn(g, ainv);
      const double thr = __dmul_rn(__dmul_rn(ainv, c.lam), c.sw[i]);
      const double v = __dsub_rn(old, t);
      const double nw = cd_shrink(v, thr);
      const double h = nw - old;
      const unsigned char nz = (unsigned char)(v != 0.0); // `x[k] -= b*a` appends iff the value is non-zero
      c.s_vnz[i] = nz;
      if (!nz) tzseen = 1;
      if (h != 0.0 && pj[u] < o.best) {
        o.best = pj[u];
        o.bk = lo + i;
        o.bh = h;
        o.bnw = nw;
      }
    };
    const long long q1 = c.pfp ? clock64() : 0;
    for (int t0 = 0; t0 < tend; t0 += NQ) {
      double x[NQ][NE], hq[NQ];
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const bool on = t0 + q < tend;
        hq[q] = on ? rh[t0 + q] : 0.0;
        const int k = on ? rk[t0 + q] : rk[t0];
        const double *col = c.A + (long long)(c.colslot ? __ldg(c.colslot + k) : k) * c.lda + lo + base + tid;
#pragma unroll
        for (int u = 0; u < NE; ++u)
          x[q][u] = (hq[q] != 0.0 && base + tid + u * COV_T < len) ? __ldg(col + u * COV_T) : 0.0;
      }
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
#pragma unroll
        for (int u = 0; u < NE; ++u) {
          if (hq[q] != 0.0) acc[u] = __dadd_rn(acc[u], __dmul_rn(x[q][u], hq[q]));
          if (TEST) accv[u] = (t0 + q < ti[u]) ? acc[u] : accv[u]; // steps 0 .. ti-1 precede the visit
        }
      }
    }
    const long long q2 = c.pfp ? clock64() : 0;
    if (TEST) {
#pragma unroll
      for (int u = 0; u < NE; ++u)
        if (ti[u] >= 0) test(u); // uniform: every lane tests its elements once, after the sweep
    }
    const long long q3 = c.pfp ? clock64() : 0;
    if (c.kx >= 0) {
      const double *col = c.A + (long long)(c.colslot ? __ldg(c.colslot + c.kx) : c.kx) * c.lda + lo + base + tid;
#pragma unroll
      for (int u = 0; u < NE; ++u)
        if (base + tid + u * COV_T < len) acc[u] = __dadd_rn(acc[u], __dmul_rn(__ldg(col + u * COV_T), c.hx));
    }
#pragma unroll
    for (int u = 0; u < NE; ++u)
      if (base + tid + u * COV_T < len) sAx2[base + tid + u * COV_T] = acc[u];
    if (c.pfp && tid == 0) {
      c.pfp[0] += q1 - q0;
      c.pfp[1] += q2 - q1;
      c.pfp[2] += q3 - q2;
      c.pfp[3] += clock64() - q3;
    }
  }
  if (TEST && tzseen) *c.tz = 1;
  return o;
}

struct PassCarry { // a full pass in progress: what a pause (lazy columns) has to carry into the next launch
  long long curpos;
  int m_old;
  double maxH;
  bool resume, paused;
  int need_k;
};

// Returns max|h| of the pass.  m_old = list length at pass start (all CTAs).  pc.resume: continue a pass that an earlier
// launch left at pc.curpos (membership flags, sorted member list and the appended entries are still in place).
template <bool PROF>
__device__ __forceinline__ double chain_pass(Ctx &c, double lam, unsigned long long pass_counter, unsigned &round,
                                             long long &accepted, bool first_pass_of_kernel, int &nonapp_total,
                                             long long *pf, int &m_bound, PassCarry &pc) {
  const CovArgs &a = c.a;
  Smem *sm = c.sm;
  const int tid = threadIdx.x;
  const bool ordered = a.randomize == 0;
  const PermKey pk = cd_perm_key((uint32_t)a.p, a.seed, pass_counter);
  const int lo = c.lo, len = c.len, rank = c.rank;
  const int m_old = pc.m_old;
  int *sorted_k = a.iscr + 7 * (long long)a.p;                                 // members in visit order
  unsigned *sorted_pos = reinterpret_cast<unsigned *>(a.iscr + 6 * (long long)a.p); // ... and their visit positions
  long long curpos = pc.resume ? pc.curpos : -1;
  double maxH = pc.resume ? pc.maxH : 0.0;
  pc.paused = false;
  if (!pc.resume) {
    for (int i = tid; i < len; i += COV_T) {
      c.s_in[i] = first_pass_of_kernel ? a.inlist[lo + i] : (unsigned char)(c.sbeta[i] != 0.0);
      c.s_vnz[i] = 1;
    }
    if (tid == 0) sm->tz = 0;
    if (rank == 0) { // sort the members by visit position (rank by counting; positions are distinct)
      for (int e = tid; e < m_old; e += COV_T) c.e_vpos[e] = visit_pos(ordered, pk, a.act[e]);
      __syncthreads();
      for (int e = tid; e < m_old; e += COV_T) {
        const unsigned my = c.e_vpos[e];
        int r = 0;
        for (int j = 0; j < m_old; ++j) r += c.e_vpos[j] < my;
        sorted_k[r] = a.act[e];
        sorted_pos[r] = my;
      }
    }
    __syncthreads();
  }
  for (;;) {
    const long long ta = PROF ? clock64() : 0;
    // ---- (1) CTA 0: the chain over the members still to be visited, in visit order
    if (rank == 0) {
      __syncthreads(); // sorted_k / sorted_pos written by this CTA's threads above are read below
      int s0 = 0;      // members already visited: those with position <= curpos
      {
        int l = 0, r = m_old;
        while (l < r) {
          const int mid = (l + r) >> 1;
          if ((long long)__ldcg(sorted_pos + mid) <= curpos) l = mid + 1; else r = mid;
        }
        s0 = l;
      }
      const int mR = m_old - s0;
      for (int t = tid; t < mR; t += COV_T) {
        const int k = __ldcg(sorted_k + s0 + t);
        c.s_act[t] = k;
        c.e_vpos[t] = __ldcg(sorted_pos + s0 + t);
        c.e_be[t] = slice_get(c, c.sbeta, k);
        c.e_g[t] = slice_get(c, c.sAx, k);
        c.e_h[t] = 0.0;
      }
      if (tid == 0) sm->mR = mR;
      __syncthreads();
      if (mR > 0) {
        chain::State S;
        S.m = mR;
        S.row = c.s_act;
        S.coord = c.s_act;
        S.g = c.e_g;
        S.be = c.e_be;
        S.ord = c.e_ord;
        S.pos = c.e_pos;
        S.stage = c.e_stage;
        S.sh = &sm->ch;
        S.G = a.A;
        S.ldg = a.lda;
        S.slot = a.colslot;
        S.prof = nullptr;
        S.hout = c.e_h;
        const CovPolicy P{a.b, a.ainv, a.omega, lam};
        (void)chain::run<COV_T>(S, P, 0.0, 1, pass_counter, true, a.seed, a.optTol, nullptr);
      }
    }
    c.cluster.sync(); // the chain's entries (k, pos, h, new value) stand in CTA 0's shared memory
    const long long tb = PROF ? clock64() : 0;
    const int mR = *c.cluster.map_shared_rank(&sm->mR, 0);
    if (rank != 0) {
      const int *rk0 = c.cluster.map_shared_rank(c.s_act, 0);
      const unsigned *rp0 = c.cluster.map_shared_rank(c.e_vpos, 0);
      const double *rh0 = c.cluster.map_shared_rank(c.e_h, 0), *rb0 = c.cluster.map_shared_rank(c.e_be, 0);
      for (int t = tid; t < mR; t += COV_T) {
        c.s_act[t] = rk0[t];
        c.e_vpos[t] = rp0[t];
        c.e_h[t] = rh0[t];
        c.e_be[t] = rb0[t];
      }
    }
    __syncthreads();
    // ---- (2) verification sweep + (3) election of the first entering coordinate
    SweepIn si;
    si.A = a.A;
    si.lda = a.lda;
    si.colslot = a.colslot;
    si.sAx = c.sAx;
    si.sb = c.sb;
    si.sainv = c.sainv;
    si.sw = c.sw;
    si.sbeta = c.sbeta;
    si.sAx2 = c.sAx2;
    si.rk = c.s_act;
    si.rh = c.e_h;
    si.rpos = c.e_vpos;
    si.s_in = c.s_in;
    si.s_vnz = c.s_vnz;
    si.tz = &sm->tz;
    si.lo = lo;
    si.len = len;
    si.tend = mR;
    si.kx = -1;
    si.ordered = ordered ? 1 : 0;
    si.lam = lam;
    si.hx = 0.0;
    si.curpos = curpos;
    si.pk = pk;
    si.pfp = (PROF && a.prof && rank == 0) ? a.prof + 10 : nullptr;
    const SweepOut so = sweep_members<true>(si);
    __syncthreads(); // sm->tz
    int tzsum = 0;
    const long long tc = PROF ? clock64() : 0;
    const Cand w = elect(c, round, so.best, so.bk, so.bh, so.bnw, sm->tz, tzsum);
    const long long td = PROF ? clock64() : 0;
    if (PROF) pf[7] += tb - ta;
    if (PROF) pf[9] += tc - tb;
    if (PROF) pf[8] += td - tc;
    int tend = mR;
    if (w.key != KEY_NONE) { // chain entries visited before the entering coordinate stay valid
      int l = 0, r = mR;
      while (l < r) {
        const int mid = (l + r) >> 1;
        if (c.e_vpos[mid] < w.key) l = mid + 1; else r = mid;
      }
      tend = l;
      if (a.colslot && __ldg(a.colslot + w.k) < 0) { // its column has not been formed: leave before applying anything
        pc.paused = true;
        pc.need_k = w.k;
        pc.curpos = curpos;
        pc.maxH = maxH;
        return maxH;
      }
      si.tend = tend;
      si.kx = w.k;
      si.hx = w.h;
      (void)sweep_members<false>(si);
    }
    // ---- commit: Ax <- Ax2 (same pointer swap in every CTA), new values of the visited members, the entering step
    {
      double *t = c.sAx;
      c.sAx = c.sAx2;
      c.sAx2 = t;
    }
    long long nacc = 0;
    for (int t = tid; t < tend; t += COV_T) {
      const double h = c.e_h[t];
      if (h != 0.0) {
        nacc += 1;
        maxH = fmax(maxH, fabs(h));
        const int k = c.s_act[t];
        if (k >= lo && k < lo + len) c.sbeta[k - lo] = c.e_be[t];
      }
    }
    if (w.key != KEY_NONE) {
      const int k = w.k;
      if (k >= lo && k < lo + len && tid == 0) {
        c.sbeta[k - lo] = w.nw;
        c.s_in[k - lo] = 1;
      }
      if (rank == 0 && tid == 0) { // setindex! appends on the first non-zero store
        a.act[sm->nact] = k;
        sm->nact += 1;
      }
    }
    // block-wide: accepted count and max|h| of the committed chain entries (every CTA computes the same numbers)
    {
      for (int o = 16; o > 0; o >>= 1) {
        nacc += __shfl_xor_sync(0xffffffffu, nacc, o);
        maxH = fmax(maxH, __shfl_xor_sync(0xffffffffu, maxH, o));
      }
      double *redd = reinterpret_cast<double *>(sm->wcand); // 16 x 32 bytes: free between elections
      if ((tid & 31) == 0) {
        redd[2 * (tid >> 5)] = (double)nacc;
        redd[2 * (tid >> 5) + 1] = maxH;
      }
      __syncthreads();
      double na = 0.0;
      for (int wq = 0; wq < COV_T / 32; ++wq) {
        na += redd[2 * wq];
        maxH = fmax(maxH, redd[2 * wq + 1]);
      }
      accepted += (long long)na;
    }
    c.cluster.sync(); // every slice is committed before anybody (the next chain, the list update) reads it remotely
    if (w.key == KEY_NONE) {
      nonapp_total = tzsum;
      break;
    }
    m_bound += 1;
    maxH = fmax(maxH, fabs(w.h));
    accepted += 1;
    curpos = (long long)w.key;
  }
  return maxH;
}

// ------------------------------------------------------------------ full pass with safe screening --
// (round 2) The verification sweep above reads |members| x p/C doubles per CTA in every round, and so does the refresh
// after an active phase: the part of the kernel that is bound by the cluster's path to L2.  Almost all of it only
// confirms, again and again, that a far-from-threshold coordinate still does not move.  With
//     S_j = sum_{k in D} A[j,k]^2        D = every coordinate whose beta changed since the last resync
// Cauchy-Schwarz gives |Ax_j(now) - Ax_j(sync)| <= sqrt(S_j) * ||beta - beta_sync||_2 for ANY matrix A, so a
// non-member row whose gradient at the last resync stays below its threshold by more than that bound provably does
// not move and is not touched at all (STALE row: its stored Ax is the value at the last resync).  Rows that fail the
// test are promoted to CURRENT: their Ax is brought up to date once (sum over D) and from then on they receive every
// committed step in order, exactly as before, and are tested exactly where their visit falls.  Members are always
// CURRENT.  When too many rows are CURRENT, and at the end of a launch, a RESYNC brings every row up to date with one
// GEMV over the columns of D and restarts the bookkeeping.  Decisions are exact; the values of Ax differ from the
// in-order accumulation only in rounding (as after a refresh).
constexpr int SC_NCMAX = 384; // CURRENT rows per CTA that trigger a resync

struct ScreenCtx {
  double *sS, *sbsync;      // per slice: S_j; beta at the last resync
  unsigned char *s_cur;     // per slice: row is CURRENT
  unsigned short *cur_list; // [L] slice indices of the CURRENT rows
  unsigned char *dflag;     // global [p]: coordinate is in D
  int *dk;                  // global: coordinates of D
  double *dsync, *dcur;     // global: beta at the last resync / beta now, per D entry
  int mode;                 // 1: normal; 2 (diagnostic): every tested row is promoted (the bound is never trusted)
  int ncmax;                // CURRENT rows per CTA that trigger a resync
};

// value of a per-slice array element through the owner CTA
__device__ __forceinline__ double slice_get_d(const Ctx &c, double *arr_local, int k) { return slice_get(c, arr_local, k); }

// every row exact again; D = current members; S, flags and the CURRENT list rebuilt.  Cluster-collective.
// Members = CTA 0's current list.
__device__ void screen_resync(Ctx &c, ScreenCtx &sc, bool first) {
  const CovArgs &a = c.a;
  Smem *sm = c.sm;
  const int tid = threadIdx.x;
  if (a.prof && c.rank == 0 && tid == 0) a.prof[13] += 1;
  if (first)
    for (int j = tid; j < c.len; j += COV_T) sc.dflag[c.lo + j] = 0;
  __threadfence();
  c.cluster.sync(); // CTA 0's list is final
  const int m = *c.cluster.map_shared_rank(&sm->nact, 0);
  // (1) CTA 0: current beta of every D entry
  int nd = first ? 0 : *c.cluster.map_shared_rank(&sm->nd, 0);
  if (c.rank == 0 && !first) {
    for (int q = tid; q < nd; q += COV_T) sc.dcur[q] = slice_get(c, c.sbeta, sc.dk[q]);
    __threadfence();
  }
  c.cluster.sync();
  // (2) all: STALE rows += sum over D of A[j,k] (beta_k - beta_k_sync)
  if (!first && nd > 0) {
    for (int j = tid; j < c.len; j += COV_T) {
      if (sc.s_cur[j]) continue;
      double acc0 = 0.0, acc1 = 0.0;
      int q = 0;
      for (; q + 2 <= nd; q += 2) {
        const int k0 = __ldcg(sc.dk + q), k1 = __ldcg(sc.dk + q + 1);
        const double d0 = __ldcg(sc.dcur + q) - __ldcg(sc.dsync + q), d1 = __ldcg(sc.dcur + q + 1) - __ldcg(sc.dsync + q + 1);
        const double v0 = d0 != 0.0 ? __ldg(col_ptr(a, k0) + c.lo + j) : 0.0;
        const double v1 = d1 != 0.0 ? __ldg(col_ptr(a, k1) + c.lo + j) : 0.0;
        acc0 = fma(v0, d0, acc0);
        acc1 = fma(v1, d1, acc1);
      }
      for (; q < nd; ++q) {
        const double d0 = __ldcg(sc.dcur + q) - __ldcg(sc.dsync + q);
        if (d0 != 0.0) acc0 = fma(__ldg(col_ptr(a, __ldcg(sc.dk + q)) + c.lo + j), d0, acc0);
      }
      c.sAx[j] += acc0 + acc1;
    }
  }
  c.cluster.sync(); // everybody is done with the old D
  // (3) new bookkeeping: D = members, S_j over the members' columns, CURRENT = members
  if (c.rank == 0) {
    for (int q = tid; q < nd; q += COV_T) sc.dflag[sc.dk[q]] = 0;
    __syncthreads();
    for (int e = tid; e < m; e += COV_T) sc.dflag[a.act[e]] = 1;
    for (int e = tid; e < m; e += COV_T) {
      const int k = a.act[e];
      sc.dk[e] = k;
      sc.dsync[e] = slice_get(c, c.sbeta, k);
    }
    if (tid == 0) sm->nd = m;
    __threadfence();
  }
  for (int j = tid; j < c.len; j += COV_T) {
    sc.s_cur[j] = 0;
    sc.sbsync[j] = c.sbeta[j];
  }
  if (tid == 0) sm->nc = 0;
  c.cluster.sync();
  for (int j = tid; j < c.len; j += COV_T) {
    double s0 = 0.0, s1 = 0.0;
    int e = 0;
    for (; e + 2 <= m; e += 2) {
      const double v0 = __ldg(col_ptr(a, __ldcg(sc.dk + e)) + c.lo + j), v1 = __ldg(col_ptr(a, __ldcg(sc.dk + e + 1)) + c.lo + j);
      s0 = fma(v0, v0, s0);
      s1 = fma(v1, v1, s1);
    }
    for (; e < m; ++e) {
      const double v0 = __ldg(col_ptr(a, __ldcg(sc.dk + e)) + c.lo + j);
      s0 = fma(v0, v0, s0);
    }
    sc.sS[j] = s0 + s1;
  }
  __syncthreads();
  for (int e = tid; e < m; e += COV_T) { // members are CURRENT
    const int k = __ldcg(sc.dk + e);
    if (k >= c.lo && k < c.lo + c.len) {
      sc.s_cur[k - c.lo] = 1;
      sc.cur_list[atomicAdd(&sm->nc, 1)] = (unsigned short)(k - c.lo);
    }
  }
  __syncthreads();
}

// after an active phase: the change of beta reaches the CURRENT rows only (everybody else is covered by the bound)
__device__ void refresh_cur(Ctx &c, ScreenCtx &sc, int m0) {
  const CovArgs &a = c.a;
  const int tid = threadIdx.x;
  const double *dlt = a.scr + a.p;
  const int *act0 = a.iscr, *act0c = a.iscr + a.p;
  const int nc = c.sm->nc;
  for (int r = tid; r < nc; r += COV_T) {
    const int j = sc.cur_list[r];
    const double *row = a.A + c.lo + j;
    double acc0 = 0.0, acc1 = 0.0;
    int i = 0;
    for (; i + 2 <= m0; i += 2) {
      const double d0 = __ldcg(dlt + i), d1 = __ldcg(dlt + i + 1);
      const double v0 = d0 != 0.0 ? __ldg(row + (long long)__ldcg(act0c + i) * a.lda) : 0.0;
      const double v1 = d1 != 0.0 ? __ldg(row + (long long)__ldcg(act0c + i + 1) * a.lda) : 0.0;
      acc0 = fma(v0, d0, acc0);
      acc1 = fma(v1, d1, acc1);
    }
    for (; i < m0; ++i) {
      const double d0 = __ldcg(dlt + i);
      if (d0 != 0.0) acc0 = fma(__ldg(row + (long long)__ldcg(act0c + i) * a.lda), d0, acc0);
    }
    c.sAx[j] += acc0 + acc1;
  }
  for (int i = tid; i < m0; i += COV_T) {
    const int k = __ldcg(act0 + i);
    if (k >= c.lo && k < c.lo + c.len) c.sbeta[k - c.lo] = __ldcg(a.beta + k);
  }
  __syncthreads();
}

// The full pass of chain_pass with screened verification.  Same contract; additionally returns (through need_resync)
// whether some CTA holds too many CURRENT rows.
template <bool PROF>
__device__ __forceinline__ double chain_pass_sc(Ctx &c, ScreenCtx &sc, double lam, unsigned long long pass_counter,
                                                unsigned &round, long long &accepted, bool first_pass_of_kernel,
                                                int &nonapp_total, long long *pf, int &m_bound, PassCarry &pc,
                                                bool &need_resync) {
  const CovArgs &a = c.a;
  Smem *sm = c.sm;
  const int tid = threadIdx.x;
  const bool ordered = a.randomize == 0;
  const PermKey pk = cd_perm_key((uint32_t)a.p, a.seed, pass_counter);
  const int lo = c.lo, len = c.len, rank = c.rank;
  const int m_old = pc.m_old;
  int *sorted_k = a.iscr + 7 * (long long)a.p;
  unsigned *sorted_pos = reinterpret_cast<unsigned *>(a.iscr + 6 * (long long)a.p);
  long long curpos = pc.resume ? pc.curpos : -1;
  double maxH = pc.resume ? pc.maxH : 0.0;
  pc.paused = false;
  need_resync = false;
  if (!pc.resume) {
    for (int i = tid; i < len; i += COV_T) {
      c.s_in[i] = first_pass_of_kernel ? a.inlist[lo + i] : (unsigned char)(c.sbeta[i] != 0.0);
      c.s_vnz[i] = 1;
    }
    if (tid == 0) sm->tz = 0;
    if (rank == 0) {
      for (int e = tid; e < m_old; e += COV_T) c.e_vpos[e] = visit_pos(ordered, pk, a.act[e]);
      __syncthreads();
      for (int e = tid; e < m_old; e += COV_T) {
        const unsigned my = c.e_vpos[e];
        int r = 0;
        for (int j = 0; j < m_old; ++j) r += c.e_vpos[j] < my;
        sorted_k[r] = a.act[e];
        sorted_pos[r] = my;
      }
    }
    __syncthreads();
  }
  int s0 = 0; // members already visited in this pass
  if (pc.resume && rank == 0) {
    int l = 0, r = m_old;
    while (l < r) {
      const int mid = (l + r) >> 1;
      if ((long long)sorted_pos[mid] <= curpos) l = mid + 1; else r = mid;
    }
    s0 = l;
  }
  for (;;) {
    const long long ta = PROF ? clock64() : 0;
    // ---- (1) CTA 0: chain over the members still to be visited; the norm bound of this round
    if (rank == 0) {
      __syncthreads();
      const int mR = m_old - s0;
      for (int t = tid; t < mR; t += COV_T) {
        const int k = sorted_k[s0 + t];
        c.s_act[t] = k;
        c.e_vpos[t] = sorted_pos[s0 + t];
        c.e_be[t] = slice_get(c, c.sbeta, k);
        c.e_g[t] = slice_get(c, c.sAx, k);
        c.e_h[t] = 0.0;
      }
      if (tid == 0) sm->mR = mR;
      // beta now of every D entry (for catch-ups) and the squared distance from the last resync
      const int nd = sm->nd;
      double part = 0.0;
      for (int q = tid; q < nd; q += COV_T) {
        const double cur = slice_get(c, c.sbeta, sc.dk[q]);
        sc.dcur[q] = cur;
        const double d = cur - sc.dsync[q];
        part = fma(d, d, part);
      }
      __syncthreads();
      if (mR > 0) {
        chain::State S;
        S.m = mR;
        S.row = c.s_act;
        S.coord = c.s_act;
        S.g = c.e_g;
        S.be = c.e_be;
        S.ord = c.e_ord;
        S.pos = c.e_pos;
        S.stage = c.e_stage;
        S.sh = &sm->ch;
        S.G = a.A;
        S.ldg = a.lda;
        S.slot = a.colslot;
        S.prof = nullptr;
        S.hout = c.e_h;
        const CovPolicy P{a.b, a.ainv, a.omega, lam};
        (void)chain::run<COV_T>(S, P, 0.0, 1, pass_counter, true, a.seed, a.optTol, nullptr);
      }
      // entries that moved are further from (or closer to) beta_sync: take the larger of the two distances
      for (int t = tid; t < mR; t += COV_T) {
        const double h = c.e_h[t];
        if (h != 0.0) {
          const double bs = slice_get(c, sc.sbsync, c.s_act[t]);
          const double dn = c.e_be[t] - bs, dold = (c.e_be[t] - h) - bs;
          part += fmax(0.0, dn * dn - dold * dold);
        }
      }
      part = warp_sum(part);
      double *redd = reinterpret_cast<double *>(sm->wcand);
      if ((tid & 31) == 0) redd[tid >> 5] = part;
      __syncthreads();
      if (tid == 0) {
        double tot = 0.0;
        for (int wq = 0; wq < COV_T / 32; ++wq) tot += redd[wq];
        sm->Dn2 = tot * (1.0 + 1e-9);
      }
      __threadfence();
      __syncthreads();
    }
    c.cluster.sync();
    const long long tb = PROF ? clock64() : 0;
    const int mR = *c.cluster.map_shared_rank(&sm->mR, 0);
    const int nd = *c.cluster.map_shared_rank(&sm->nd, 0);
    const double Dn2 = *c.cluster.map_shared_rank(&sm->Dn2, 0);
    if (rank != 0) {
      const int *rk0 = c.cluster.map_shared_rank(c.s_act, 0);
      const unsigned *rp0 = c.cluster.map_shared_rank(c.e_vpos, 0);
      const double *rh0 = c.cluster.map_shared_rank(c.e_h, 0), *rb0 = c.cluster.map_shared_rank(c.e_be, 0);
      for (int t = tid; t < mR; t += COV_T) {
        c.s_act[t] = rk0[t];
        c.e_vpos[t] = rp0[t];
        c.e_h[t] = rh0[t];
        c.e_be[t] = rb0[t];
      }
    }
    __syncthreads();
    // ---- (2a) STALE rows: safe-screening test; the few that fail are brought up to date and become CURRENT
    const long long u0 = PROF ? clock64() : 0;
    const int nc_old = sm->nc;
    __syncthreads();
    for (int i = tid; i < len; i += COV_T) {
      if (sc.s_cur[i] || c.s_in[i]) continue;
      const unsigned pj = visit_pos(ordered, pk, lo + i);
      if ((long long)pj <= curpos) continue;
      const double g = c.sAx[i] + c.sb[i];
      const double d = lam * c.sw[i] * (1.0 - 1e-9) - fabs(g);
      if (sc.mode == 1 && d > 0.0 && d * d > sc.sS[i] * Dn2) continue; // provably below the threshold whenever its turn comes
      sc.s_cur[i] = 1;
      sc.cur_list[atomicAdd(&sm->nc, 1)] = (unsigned short)i;
    }
    __syncthreads();
    const long long u1 = PROF ? clock64() : 0;
    // catch-up of the rows just promoted, one warp per row: Ax_j += sum over D of A[j,k] (beta_k - beta_k at the resync)
    {
      const int nc_new = sm->nc, lane = tid & 31, warp = tid >> 5;
      if (nc_new > nc_old) { // (uniform per CTA) D goes to shared memory first: one trip to L2 instead of three per term
        double *dd_s = c.e_g;                              // free between the chain and the next round (nd <= ecap)
        int *dk_s = reinterpret_cast<int *>(c.e_stage);    // the engine's stage buffers are idle here
        for (int q = tid; q < nd; q += COV_T) {
          dd_s[q] = __ldcg(sc.dcur + q) - __ldcg(sc.dsync + q);
          dk_s[q] = __ldcg(sc.dk + q);
        }
        __syncthreads();
        for (int r = nc_old + warp; r < nc_new; r += COV_T / 32) {
          const int i = sc.cur_list[r];
          double acc0 = 0.0, acc1 = 0.0;
          int q = lane;
          for (; q + 32 < nd; q += 64) {
            const double d0 = dd_s[q], d1 = dd_s[q + 32];
            const double v0 = d0 != 0.0 ? __ldg(col_ptr(a, dk_s[q]) + lo + i) : 0.0;
            const double v1 = d1 != 0.0 ? __ldg(col_ptr(a, dk_s[q + 32]) + lo + i) : 0.0;
            acc0 = fma(v0, d0, acc0);
            acc1 = fma(v1, d1, acc1);
          }
          if (q < nd) {
            const double d0 = dd_s[q];
            if (d0 != 0.0) acc0 = fma(__ldg(col_ptr(a, dk_s[q]) + lo + i), d0, acc0);
          }
          acc0 = warp_sum(acc0 + acc1);
          if (lane == 0) c.sAx[i] += acc0;
        }
      }
      if (PROF && rank == 0 && tid == 0) {
        a.prof[10] += nc_new - nc_old;
        a.prof[11] += nc_new;
        a.prof[12] += 1;
      }
    }
    __syncthreads();
    const long long u2 = PROF ? clock64() : 0;
    // ---- (2b) CURRENT rows: the chain steps in visit order, exact test where the row's own visit falls
    const int nc = sm->nc;
    unsigned best = KEY_NONE;
    int bk = 0;
    double bh = 0.0, bnw = 0.0;
    int tzseen = 0;
    for (int r = tid; r < nc; r += COV_T) {
      const int i = sc.cur_list[r];
      double acc = c.sAx[i], accv = acc;
      const bool testme = !c.s_in[i];
      const unsigned pj = visit_pos(ordered, pk, lo + i);
      const bool pend = testme && (long long)pj > curpos;
      for (int t0 = 0; t0 < mR; t0 += 8) {
        double x[8], hq[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const bool on = t0 + q < mR;
          hq[q] = on ? c.e_h[t0 + q] : 0.0;
          x[q] = hq[q] != 0.0 ? __ldg(col_ptr(a, c.s_act[t0 + q]) + lo + i) : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (hq[q] != 0.0) acc = __dadd_rn(acc, __dmul_rn(x[q], hq[q]));
          if (t0 + q < mR && c.e_vpos[t0 + q] < pj) accv = acc; // steps visited before this row
        }
      }
      c.sAx2[i] = acc;
      if (pend) {
        const double ainv = c.sainv[i];
        const double g = accv + c.sb[i];
        const double old = c.sbeta[i];
        const double t = __dmul_rn(g, ainv);
        const double thr = __dmul_rn(__dmul_rn(ainv, lam), c.sw[i]);
        const double v = __dsub_rn(old, t);
        const double nw = cd_shrink(v, thr);
        const double h = nw - old;
        const unsigned char nz = (unsigned char)(v != 0.0);
        c.s_vnz[i] = nz;
        if (!nz) tzseen = 1;
        if (h != 0.0 && pj < best) {
          best = pj;
          bk = lo + i;
          bh = h;
          bnw = nw;
        }
      }
    }
    if (tzseen) sm->tz = 1;
    __syncthreads();
    int tzsum = 0;
    const long long tc = PROF ? clock64() : 0;
    if (PROF && rank == 0 && tid == 0) {
      a.prof[14] += u0 - tb;
      a.prof[15] += u1 - u0;
      a.prof[23] += u2 - u1;
      a.prof[3 + 20] += 0;
      a.prof[22 + 1] += 0;
    }
    if (PROF && rank == 0 && tid == 0) a.prof[9 + 5] += 0;
    const long long u3 = tc - u2;
    if (PROF && rank == 0 && tid == 0) a.prof[8 + 0] += 0 * u3;
    const Cand w = elect(c, round, best, bk, bh, bnw, sm->tz | ((nc > sc.ncmax || nd + 64 > c.ecap) ? 0x10000 : 0), tzsum);
    const long long td = PROF ? clock64() : 0;
    if (PROF) pf[7] += tb - ta;
    if (PROF) pf[9] += tc - tb;
    if (PROF) pf[8] += td - tc;
    if (tzsum >= 0x10000) need_resync = true;
    tzsum &= 0xffff;
    int tend = mR;
    if (w.key != KEY_NONE) {
      int l = 0, r = mR;
      while (l < r) {
        const int mid = (l + r) >> 1;
        if (c.e_vpos[mid] < w.key) l = mid + 1; else r = mid;
      }
      tend = l;
      if (a.colslot && __ldg(a.colslot + w.k) < 0) { // column not formed: leave before applying anything of this round
        pc.paused = true;
        pc.need_k = w.k;
        pc.curpos = curpos;
        pc.maxH = maxH;
        return maxH;
      }
    }
    // ---- commit on the CURRENT rows: all steps, or the steps before the entering coordinate plus its own
    if (w.key == KEY_NONE) {
      for (int r = tid; r < nc; r += COV_T) {
        const int i = sc.cur_list[r];
        c.sAx[i] = c.sAx2[i];
      }
    } else {
      const double *colx = col_ptr(a, w.k) + lo;
      for (int r = tid; r < nc; r += COV_T) {
        const int i = sc.cur_list[r];
        double acc = c.sAx[i];
        for (int t = 0; t < tend; ++t) {
          const double h = c.e_h[t];
          if (h != 0.0) acc = __dadd_rn(acc, __dmul_rn(__ldg(col_ptr(a, c.s_act[t]) + lo + i), h));
        }
        c.sAx[i] = __dadd_rn(acc, __dmul_rn(__ldg(colx + i), w.h));
      }
      // the entering coordinate joins D: S of every row grows by its column (one coalesced read per entering step)
      for (int i = tid; i < len; i += COV_T) {
        const double v = __ldg(colx + i);
        sc.sS[i] = fma(v, v, sc.sS[i]);
      }
    }
    long long nacc = 0;
    for (int t = tid; t < tend; t += COV_T) {
      const double h = c.e_h[t];
      if (h != 0.0) {
        nacc += 1;
        maxH = fmax(maxH, fabs(h));
        const int k = c.s_act[t];
        if (k >= lo && k < lo + len) c.sbeta[k - lo] = c.e_be[t];
      }
    }
    if (w.key != KEY_NONE) {
      const int k = w.k;
      if (k >= lo && k < lo + len && tid == 0) {
        c.sbeta[k - lo] = w.nw;
        c.s_in[k - lo] = 1; // (its row was promoted to CURRENT by the screening test: it moved)
      }
      if (rank == 0 && tid == 0) {
        a.act[sm->nact] = k;
        sm->nact += 1;
        if (!sc.dflag[k]) { // (a coordinate that left and comes back is still in D with its beta_sync)
          sc.dflag[k] = 1;
          sc.dk[sm->nd] = k; // beta_sync of a coordinate that was not a member is 0
          sc.dsync[sm->nd] = 0.0;
          sm->nd += 1;
        }
        __threadfence();
      }
    }
    {
      for (int o = 16; o > 0; o >>= 1) {
        nacc += __shfl_xor_sync(0xffffffffu, nacc, o);
        maxH = fmax(maxH, __shfl_xor_sync(0xffffffffu, maxH, o));
      }
      double *redd = reinterpret_cast<double *>(sm->wcand);
      __syncthreads();
      if ((tid & 31) == 0) {
        redd[2 * (tid >> 5)] = (double)nacc;
        redd[2 * (tid >> 5) + 1] = maxH;
      }
      __syncthreads();
      double na = 0.0;
      for (int wq = 0; wq < COV_T / 32; ++wq) {
        na += redd[2 * wq];
        maxH = fmax(maxH, redd[2 * wq + 1]);
      }
      accepted += (long long)na;
    }
    c.cluster.sync();
    if (w.key == KEY_NONE) {
      nonapp_total = tzsum;
      break;
    }
    m_bound += 1;
    maxH = fmax(maxH, fabs(w.h));
    accepted += 1;
    curpos = (long long)w.key;
    s0 += tend;
  }
  return maxH;
}

// rare: publish the visited non-members that were not appended (tentative value exactly zero)
__device__ void publish_nonapp(Ctx &c) {
  int *cnt = c.a.flag + 2, *list = c.a.iscr + 6 * (long long)c.a.p;
  for (int i = threadIdx.x; i < c.len; i += COV_T)
    if (!c.s_in[i] && !c.s_vnz[i]) list[atomicAdd(cnt, 1)] = c.lo + i;
  __threadfence();
}

// dropzeros! after a full pass, CTA 0.  Gathers the final values of all listed coordinates, places
// the new entries where the reference's temporary appends put them (see cd_compact_list) and compacts.
__device__ void list_update_full(Ctx &c, int m_old, int nonapp_total, unsigned long long pass_counter) {
  const CovArgs &a = c.a;
  Smem *sm = c.sm;
  const int tid = threadIdx.x;
  const int m = sm->nact;
  for (int i = tid; i < m; i += COV_T) a.actval[i] = slice_get(c, c.sbeta, a.act[i]);
  int *newpos = a.iscr + 7 * (long long)a.p;
  const bool ordered = a.randomize == 0;
  const PermKey pk = cd_perm_key((uint32_t)a.p, a.seed, pass_counter);
  const int *nonapp = a.iscr + 6 * (long long)a.p;
  __syncthreads(); // the sorted member list of the chain pass shares newpos' storage: everybody is done reading it
  for (int e = m_old + tid; e < m; e += COV_T) {
    const int k = a.act[e];
    const long long vis = ordered ? k : (long long)cd_perm_inv(pk, (uint32_t)k);
    int before = 0; // old members and non-appended non-members visited before k
    for (int j = 0; j < m_old; ++j) {
      const int kj = a.act[j];
      before += (ordered ? kj : (long long)cd_perm_inv(pk, (uint32_t)kj)) < vis;
    }
    for (int j = 0; j < nonapp_total; ++j) {
      const int kj = __ldcg(nonapp + j);
      before += (ordered ? kj : (long long)cd_perm_inv(pk, (uint32_t)kj)) < vis;
    }
    newpos[e - m_old] = m_old + (int)vis - before;
  }
  __syncthreads();
  cd_compact_list<COV_T>(a.act, a.actval, m_old, m, newpos, a.inlist, a.iscr + (long long)a.p, a.scr + 2 * (long long)a.p,
                         sm->s2);
  if (tid == 0) sm->nact = sm->s2[0];
  __syncthreads();
}

// CTA 0 after an active phase: final list, dense beta, per-snapshot-entry delta (for refresh_slice), summary
__device__ void engine_publish(Ctx &c, const chain::Result &r, int m0, const int *act0, const double *scr_b0, double *scr_dlt) {
  const CovArgs &a = c.a;
  Smem *sm = c.sm;
  const int tid = threadIdx.x;
  for (int i = tid; i < m0; i += COV_T) a.beta[act0[i]] = 0.0;
  __syncthreads();
  for (int i = tid; i < r.m; i += COV_T) {
    const int k = c.s_act[i];
    const double be = c.e_be[i];
    a.beta[k] = be;
    a.act[i] = k;
    a.actval[i] = be;
  }
  __syncthreads();
  for (int i = tid; i < m0; i += COV_T) scr_dlt[i] = a.beta[act0[i]] - scr_b0[i];
  if (tid == 0) {
    sm->nact = r.m;
    sm->bc.npasses = r.npasses;
    sm->bc.visits = r.visits;
    sm->bc.accepted = r.accepted;
    sm->bc.maxH = r.maxH;
    sm->bc.m0 = m0;
    sm->bc.nact = r.m;
    sm->bc.conv = r.conv;
  }
  __syncthreads();
}

__device__ void active_engine(Ctx &c, double lam, long long maxPasses, unsigned long long pass_counter) {
  const CovArgs &a = c.a;
  Smem *sm = c.sm;
  const int tid = threadIdx.x;
  const int m0 = sm->nact;
  double *scr_b0 = a.scr;        // [p] beta at entry, by snapshot index
  double *scr_dlt = a.scr + a.p; // [p] delta by snapshot index
  int *act0 = a.iscr;            // [p] snapshot of the list
  int *act0c = a.iscr + a.p;     // [p] ... and the column index (slot) of every snapshot entry
  for (int i = tid; i < m0; i += COV_T) {
    const int k = a.act[i];
    const double be = a.actval[i];
    c.s_act[i] = k;
    act0[i] = k;
    act0c[i] = a.colslot ? __ldg(a.colslot + k) : k;
    c.e_be[i] = be;
    scr_b0[i] = be;
    c.e_g[i] = slice_get(c, c.sAx, k);
  }
  __syncthreads();
  chain::State S;
  S.m = m0;
  S.row = c.s_act;
  S.coord = c.s_act;
  S.g = c.e_g;
  S.be = c.e_be;
  S.ord = c.e_ord;
  S.pos = c.e_pos;
  S.stage = c.e_stage;
  S.sh = &sm->ch;
  S.G = a.A;
  S.ldg = a.lda;
  S.slot = a.colslot;
  S.prof = a.prof ? a.prof + 16 : nullptr;
  const CovPolicy P{a.b, a.ainv, a.omega, lam};
  const chain::Result r = chain::run<COV_T>(S, P, 0.0, maxPasses, pass_counter, a.randomize == 0, a.seed, a.optTol, a.inlist);
  engine_publish(c, r, m0, act0, scr_b0, scr_dlt);
}

// The same phase with the chain engine spread over the whole cluster (chain_engine.cuh: run_multi), for active
// sets of many 32-entry blocks: every CTA applies each block's steps to the entries it owns, one cluster barrier per
// block.  Called by ALL CTAs; m is the list length read from CTA 0.
// c.multi_ok carries the smallest list length handed to the distributed engine (0: never).  Default 384: below that
// the single-CTA engine is chain-bound anyway and saves the cluster barrier + L2 round trip per block.
__device__ void active_engine_multi(Ctx &c, double lam, long long maxPasses, unsigned long long pass_counter, int m0) {
  const CovArgs &a = c.a;
  Smem *sm = c.sm;
  const int tid = threadIdx.x;
  double *scr_b0 = a.scr, *scr_dlt = a.scr + a.p;
  int *act0 = a.iscr, *act0c = a.iscr + a.p;
  double *gG = a.scr + 6 * (long long)a.p, *hG = a.scr + 8 * (long long)a.p, *pmaxG = hG + 64;
  int *flagsG = reinterpret_cast<int *>(hG + 72);
  for (int i = tid; i < m0; i += COV_T) c.s_act[i] = a.act[i]; // every CTA: its own copy of the list
  if (c.rank == 0) {
    for (int i = tid; i < m0; i += COV_T) {
      const int k = a.act[i];
      const double be = a.actval[i];
      act0[i] = k;
      act0c[i] = a.colslot ? __ldg(a.colslot + k) : k;
      c.e_be[i] = be;
      scr_b0[i] = be;
      gG[i] = slice_get(c, c.sAx, k);
    }
  }
  __syncthreads();
  c.cluster.sync(); // g published
  chain::State S;
  S.m = m0;
  S.row = c.s_act;
  S.coord = c.s_act;
  S.g = nullptr;
  S.be = c.e_be;
  S.ord = c.e_ord;
  S.pos = c.e_pos;
  S.stage = c.e_stage;
  S.sh = &sm->ch;
  S.G = a.A;
  S.ldg = a.lda;
  S.slot = a.colslot;
  S.prof = nullptr;
  chain::Multi X{c.C, c.rank, gG, hG, pmaxG, flagsG, a.act, reinterpret_cast<uint4 *>(a.chain_scr), reinterpret_cast<uint4 *>(a.chain_scr) + CD_GCAP};
  const CovPolicy P{a.b, a.ainv, a.omega, lam};
  cg::cluster_group &cl = c.cluster;
  const chain::Result r = chain::run_multi<COV_T>(S, X, P, [&cl]() { cl.sync(); }, 0.0, maxPasses, pass_counter,
                                                  a.randomize == 0, a.seed, a.optTol, a.inlist);
  if (c.rank == 0) engine_publish(c, r, m0, act0, scr_b0, scr_dlt);
}

// every CTA after the active phase: Ax[slice] += A[slice, act0] * dlt ; beta[slice] <- dense beta
__device__ void refresh_slice(Ctx &c, int m0) {
  const CovArgs &a = c.a;
  const int tid = threadIdx.x;
  const double *dlt = a.scr + a.p;
  const int *act0 = a.iscr, *act0c = a.iscr + a.p;
  for (int j = tid; j < c.len; j += COV_T) {
    const double *row = a.A + c.lo + j;
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
    int i = 0;
    for (; i + 4 <= m0; i += 4) {
      const int k0 = __ldcg(act0c + i), k1 = __ldcg(act0c + i + 1), k2 = __ldcg(act0c + i + 2), k3 = __ldcg(act0c + i + 3);
      const double d0 = __ldcg(dlt + i), d1 = __ldcg(dlt + i + 1), d2 = __ldcg(dlt + i + 2), d3 = __ldcg(dlt + i + 3);
      const double v0 = __ldg(row + (long long)k0 * a.lda), v1 = __ldg(row + (long long)k1 * a.lda);
      const double v2 = __ldg(row + (long long)k2 * a.lda), v3 = __ldg(row + (long long)k3 * a.lda);
      acc0 = fma(v0, d0, acc0);
      acc1 = fma(v1, d1, acc1);
      acc2 = fma(v2, d2, acc2);
      acc3 = fma(v3, d3, acc3);
    }
    for (; i < m0; ++i) acc0 = fma(__ldg(row + (long long)__ldcg(act0c + i) * a.lda), __ldcg(dlt + i), acc0);
    c.sAx[j] += (acc0 + acc1) + (acc2 + acc3);
  }
  for (int i = tid; i < m0; i += COV_T) {
    const int k = __ldcg(act0 + i);
    if (k >= c.lo && k < c.lo + c.len) c.sbeta[k - c.lo] = __ldcg(a.beta + k);
  }
  __syncthreads();
}

template <bool PROF>
__global__ void __launch_bounds__(COV_T, 1) cov_path_kernel(const CovArgs a, int L, int slice_in_smem, int ecap, int multi_ok,
                                                            int screen) {
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Ctx c{a, cluster};
  c.rank = (int)cluster.block_rank();
  c.C = (int)cluster.num_blocks();
  c.L = L;
  c.slice_in_smem = slice_in_smem;
  c.lo = min(c.rank * L, a.p);
  c.len = min(a.p, c.lo + L) - c.lo;
  const int tid = threadIdx.x;
  const bool resumed = a.resume != nullptr && a.resume->valid != 0;
  // carve shared memory (identical layout in every CTA)
  unsigned char *sp = smem_raw;
  c.sm = reinterpret_cast<Smem *>(sp);
  sp += (sizeof(Smem) + 15) / 16 * 16;
  c.ecap = ecap;
  c.multi_ok = multi_ok;
  c.e_stage = reinterpret_cast<double *>(sp);
  sp += chain::STAGE_DOUBLES * sizeof(double);
  c.e_g = reinterpret_cast<double *>(sp);
  sp += (size_t)ecap * sizeof(double);
  c.e_be = reinterpret_cast<double *>(sp);
  sp += (size_t)ecap * sizeof(double);
  c.e_h = reinterpret_cast<double *>(sp);
  sp += (size_t)ecap * sizeof(double);
  c.s_act = reinterpret_cast<int *>(sp);
  sp += (size_t)ecap * sizeof(int);
  c.e_vpos = reinterpret_cast<unsigned *>(sp);
  sp += (size_t)ecap * sizeof(unsigned);
  c.e_ord = reinterpret_cast<unsigned short *>(sp);
  sp += (size_t)ecap * sizeof(unsigned short);
  c.e_pos = reinterpret_cast<unsigned short *>(sp);
  sp += (size_t)ecap * sizeof(unsigned short);
  if (slice_in_smem) {
    double *d = reinterpret_cast<double *>(sp);
    c.sAx = d;
    c.sb = d + L;
    c.sainv = d + 2 * L;
    c.sw = d + 3 * L;
    c.sbeta = d + 4 * L;
    c.sAx2 = d + 5 * L;
    c.s_in = reinterpret_cast<unsigned char *>(d + (screen ? 8 : 6) * L);
    c.s_vnz = c.s_in + L;
    for (int i = tid; i < c.len; i += COV_T) {
      const int k = c.lo + i;
      c.sAx[i] = a.Ax[k];
      c.sb[i] = a.b[k];
      c.sainv[i] = a.ainv[k];
      c.sw[i] = a.omega ? a.omega[k] : 1.0;
      c.sbeta[i] = a.beta[k];
      if (resumed) {
        c.s_in[i] = a.bscr[k];
        c.s_vnz[i] = a.bscr[a.p + k];
      }
    }
  } else { // very large p: slices stay in global memory (each CTA touches only its own)
    double *g = a.scr + 4 * (long long)a.p;
    c.sAx = a.Ax + c.lo;
    c.sAx2 = a.scr + 9 * (long long)a.p + c.lo;
    c.sbeta = g + c.lo; // private dense copy so CTA 0's writes to a.beta never race
    c.sb = const_cast<double *>(a.b) + c.lo;
    c.sainv = const_cast<double *>(a.ainv) + c.lo;
    c.sw = g + a.p + c.lo;
    c.s_in = a.bscr + c.lo;
    c.s_vnz = a.bscr + a.p + c.lo;
    for (int i = tid; i < c.len; i += COV_T) {
      c.sbeta[i] = a.beta[c.lo + i];
      c.sw[i] = a.omega ? a.omega[c.lo + i] : 1.0;
    }
  }
  if (tid == 0) {
    c.sm->nact = (c.rank == 0) ? *a.nact : 0;
    c.sm->bc.status = 0;
    c.sm->tz = resumed ? a.resume->tzflag : 0;
    c.sm->mR = 0;
    mbar_init(&c.sm->mbar[0], 1);
    mbar_init(&c.sm->mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster.sync();

  ScreenCtx sc;
  sc.sS = sc.sbsync = nullptr;
  sc.s_cur = nullptr;
  sc.cur_list = nullptr;
  sc.dk = a.iscr + 8 * (long long)a.p;
  sc.dsync = a.scr + 12 * (long long)a.p;
  sc.dcur = a.scr + 13 * (long long)a.p;
  sc.dflag = a.bscr + 2 * (long long)a.p;
  sc.mode = screen == 2 ? 2 : 1;
  sc.ncmax = screen > 2 ? screen : SC_NCMAX;
  if (screen) { // only with the slices in shared memory (launcher)
    double *d = c.sAx;
    sc.sS = d + 6 * L;
    sc.sbsync = d + 7 * L;
    sc.s_cur = c.s_vnz + L;
    sc.cur_list = reinterpret_cast<unsigned short *>(sc.s_cur + L + (L & 1));
    if (tid == 0) {
      c.sm->nd = 0;
      c.sm->nc = 0;
      c.sm->Dn2 = 0.0;
    }
    __syncthreads();
    screen_resync(c, sc, true);
  }
  long long pf[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  const long long t_start = PROF ? clock64() : 0;
  unsigned round = 0; // candidate-exchange rounds so far (selects slot parity and mbarrier phase)
  int m_bound = *a.nact; // upper bound of the list length, the same in every CTA
  int m_known = m_bound; // the exact list length when every CTA knows it, else -1
  bool first_pass = !resumed;
  unsigned long long pass_counter = 0;
  DevStats st;
  st.passes = st.full_passes = st.visits = st.accepted = 0;
  st.maxH = 0.0;
  st.converged = st.outer_iters = 0;
  st.sigma = 0.0;
  int status = 0;
  long long cols_done = 0, out_off = 0;
  int li0 = 0;
  PassCarry pc;
  pc.resume = false;
  pc.paused = false;
  pc.need_k = -1;
  pc.curpos = -1;
  pc.m_old = 0;
  pc.maxH = 0.0;
  if (resumed) {
    const CovResume &R = *a.resume;
    li0 = R.li;
    st = R.st;
    pass_counter = R.pass_counter;
    m_bound = R.m_bound;
    m_known = -1;
    out_off = R.out_off;
    cols_done = R.cols_done;
    pc.resume = true;
    pc.curpos = R.curpos;
    pc.m_old = R.m_old;
    pc.maxH = R.maxH;
  }
  int *lpos = a.iscr + 6 * (long long)a.p; // restricted event passes: list position of every listed coordinate
  for (int li = li0; li < a.nlambda && status == 0; ++li) {
    const double lam = a.lambdas[li];
    const bool resume_here = resumed && li == li0;
    if (!resume_here && (li == 0 || !a.accumulate)) {
      st.passes = st.full_passes = st.visits = st.accepted = 0;
      st.maxH = 0.0;
      st.converged = 0;
      st.outer_iters = 0;
      st.sigma = 0.0;
    }
    st.converged = 0;
    bool conv = true;
    long long iter = resume_here ? a.resume->iter : 0;
    while (iter < a.maxIter || pc.resume) {
      if (conv) {
        const bool cont = pc.resume; // continue the pass an earlier launch left
        if (!cont) {
          iter += 1;
          st.passes += 1;
          st.full_passes += 1;
          st.visits += a.p;
        }
        int nonapp_total = 0;
        bool resync_after = false, reinit_after = false;
        const long long t0 = PROF ? clock64() : 0;
        const long long acc0 = st.accepted;
        const bool use_chain = cont || (!a.events_only && m_known >= 0 && m_known <= c.ecap);
        int m_old;
        double maxH;
        if (use_chain) {
          if (!cont) pc.m_old = m_known;
          m_old = pc.m_old;
          bool need_resync = false;
          if (screen)
            maxH = chain_pass_sc<PROF>(c, sc, lam, pass_counter, round, st.accepted, first_pass, nonapp_total, pf, m_bound, pc,
                                       need_resync);
          else
            maxH = chain_pass<PROF>(c, lam, pass_counter, round, st.accepted, first_pass, nonapp_total, pf, m_bound, pc);
          resync_after = need_resync;
          pc.resume = false;
          if (pc.paused) { // an entering coordinate has no column yet: leave, the host forms it and launches again
            status = a.resume ? 3 : 4;
            if (a.resume && c.rank == 0 && tid == 0) {
              CovResume &R = *a.resume;
              R.valid = 0;
              R.need_k = pc.need_k;
              R.li = li;
              R.conv = 1;
              R.m_bound = m_bound;
              R.m_old = pc.m_old;
              R.tzflag = 1; // conservative: a later launch re-checks the flags of the slices
              R.iter = iter;
              R.out_off = out_off;
              R.cols_done = cols_done;
              R.curpos = pc.curpos;
              R.pass_counter = pass_counter;
              R.maxH = pc.maxH;
              R.st = st;
            }
            break;
          }
        } else {
          if (screen) screen_resync(c, sc, false); // the event pass keeps every row exact by itself
          m_old = c.sm->nact; // only meaningful on CTA 0
          maxH = event_pass<PROF, false>(c, lam, pass_counter, round, st.accepted, first_pass, nonapp_total, pf, m_bound,
                                         nullptr, 0);
          reinit_after = screen != 0;
        }
        const long long t1 = PROF ? clock64() : 0;
        if (PROF) pf[0] += t1 - t0;
        if (PROF) pf[4] += st.accepted - acc0;
        first_pass = false;
        if (nonapp_total > 0) { // rare slow path: collect the non-appended coordinates for CTA 0
          if (c.rank == 0 && tid == 0) a.flag[2] = 0;
          cluster.sync();
          publish_nonapp(c);
          cluster.sync();
          nonapp_total = __ldcg(a.flag + 2);
        }
        if (c.rank == 0) list_update_full(c, m_old, nonapp_total, pass_counter);
        if (reinit_after) screen_resync(c, sc, true);
        else if (resync_after) screen_resync(c, sc, false);
        if (PROF) pf[1] += clock64() - t1;
        pass_counter += 1;
        m_known = -1;
        st.maxH = maxH;
        conv = maxH < a.optTol;
        if (conv) {
          st.converged = 1;
          break;
        }
      } else {
        const long long t0 = PROF ? clock64() : 0;
        int m_all = -1; // list length, fetched from CTA 0 only when it matters
        if (c.C == 1) {
          m_all = c.sm->nact;
        } else if ((c.multi_ok > 0 && m_bound >= c.multi_ok) || m_bound > c.ecap) {
          cluster.sync(); // CTA 0 has finished the list update
          m_all = *cluster.map_shared_rank(&c.sm->nact, 0);
        }
        if (m_all > c.ecap) {
          // ---- list longer than the engine holds: one event-by-event pass over the list (exact, slow, no size limit)
          if (screen) screen_resync(c, sc, false);
          if (c.rank == 0) {
            for (int e = tid; e < m_all; e += COV_T) lpos[a.act[e]] = e;
            __threadfence();
          }
          cluster.sync();
          int dummy = 0;
          const long long acc0 = st.accepted;
          const double maxH = event_pass<PROF, true>(c, lam, pass_counter, round, st.accepted, false, dummy, pf, m_bound, lpos, m_all);
          if (c.rank == 0) list_update_full(c, m_all, 0, pass_counter);
          if (screen) screen_resync(c, sc, true);
          (void)acc0;
          iter += 1;
          pass_counter += 1;
          st.passes += 1;
          st.visits += m_all;
          st.maxH = maxH;
          conv = maxH < a.optTol;
          m_bound = m_all;
          m_known = -1;
          cluster.sync(); // the list update is complete before the next pass looks at the list
          continue;
        }
        if (c.multi_ok > 0 && m_all >= c.multi_ok) {
          active_engine_multi(c, lam, a.maxIter - iter, pass_counter, m_all);
        } else if (c.rank == 0) {
          active_engine(c, lam, a.maxIter - iter, pass_counter);
        }
        cluster.sync();
        const long long t1 = PROF ? clock64() : 0;
        if (PROF) pf[2] += t1 - t0;
        const Bcast *bc = cluster.map_shared_rank(&c.sm->bc, 0);
        const Bcast b = *bc;
        if (screen)
          refresh_cur(c, sc, b.m0);
        else
          refresh_slice(c, b.m0);
        cluster.sync(); // every slice holds the refreshed Ax / beta before CTA 0 reads them remotely (next pass or phase)
        if (PROF) pf[3] += clock64() - t1;
        if (PROF) pf[5] += b.visits;
        iter += b.npasses;
        pass_counter += b.npasses;
        st.passes += b.npasses;
        st.visits += b.visits;
        st.accepted += b.accepted;
        st.maxH = b.maxH;
        conv = b.conv != 0;
        m_bound = b.nact;
        m_known = b.nact;
        // conv == 0 here means the pass budget ran out: the while condition ends the solve
      }
    }
    if (status) break;
    // ---- end of this lambda: publish nnz, write the path column / stats
    if (c.rank == 0 && tid == 0) c.sm->bc.nact = c.sm->nact;
    cluster.sync();
    const int nnz = cluster.map_shared_rank(&c.sm->bc, 0)->nact;
    m_bound = nnz;
    m_known = nnz;
    if (!a.accumulate) {
      if (c.rank == 0) {
        if (a.colptr) {
          if (out_off + nnz > a.capacity) {
            status = 1;
          } else {
            for (int i = tid; i < nnz; i += COV_T) {
              a.rowval[out_off + i] = (long long)a.act[i] + 1;
              a.nzval[out_off + i] = a.actval[i];
            }
            if (tid == 0) a.colptr[li + 1] = out_off + nnz;
          }
        }
        if (tid == 0 && a.stats) a.stats[li] = st;
      } else if (a.colptr && out_off + nnz > a.capacity) {
        status = 1;
      }
      out_off += nnz;
      if (status == 0) cols_done = li + 1;
      if (a.max_hat_s >= 0 && nnz > a.max_hat_s) break;
    } else {
      cols_done = li + 1;
    }
  }
  if (a.accumulate && c.rank == 0 && tid == 0 && a.stats && status != 3) a.stats[0] = st;
  // ---- write the state back
  if (screen) screen_resync(c, sc, false); // every row of Ax exact again
  cluster.sync();
  if (slice_in_smem) {
    for (int i = tid; i < c.len; i += COV_T) {
      a.Ax[c.lo + i] = c.sAx[i];
      a.beta[c.lo + i] = c.sbeta[i];
      if (status == 3) {
        a.bscr[c.lo + i] = c.s_in[i];
        a.bscr[a.p + c.lo + i] = c.s_vnz[i];
      }
    }
  } else {
    for (int i = tid; i < c.len; i += COV_T) {
      a.beta[c.lo + i] = c.sbeta[i];
      if (c.sAx != a.Ax + c.lo) a.Ax[c.lo + i] = c.sAx[i]; // an odd number of chain-pass commits: Ax lives in the scratch copy
    }
  }
  if (c.rank == 0 && tid == 0) {
    if (a.prof) {
      pf[6] = clock64() - t_start;
      for (int i = 0; i < 10; ++i) a.prof[i] += pf[i]; // accumulated over the launches of one solve (host zeroes)
    }
    *a.nact = c.sm->nact;
    a.flag[0] = status;
    a.flag[1] = (int)cols_done;
  }
  cluster.sync(); // nobody exits while a peer may still read its shared memory
}

// ------------------------------------------------------------ small kernels --
// initialize!(f::CDQuadraticLoss, x): Ax = sum_i A[:, act_i] * val_i (cd_differentiable_function.jl:311-320),
// plus the dense copy of the iterate and the membership flags.
__global__ void cov_init_kernel(const double *A, long long lda, int p, const int *act, const double *actval,
                                const int *nact, double *Ax, double *beta, unsigned char *inlist, const int *colslot) {
  const int m = *nact;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < p; j += gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (int i = 0; i < m; ++i) acc += A[j + (long long)(colslot ? colslot[act[i]] : act[i]) * lda] * actval[i];
    Ax[j] = acc;
    beta[j] = 0.0;
    inlist[j] = 0;
  }
}
__global__ void scatter_iterate_kernel(const int *act, const double *actval, const int *nact, double *beta,
                                       unsigned char *inlist) {
  const int m = *nact;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    beta[act[i]] = actval[i];
    inlist[act[i]] = 1;
  }
}
__global__ void diag_sqrt_kernel(const double *A, long long lda, int p, double *out) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < p; j += gridDim.x * blockDim.x)
    out[j] = sqrt(A[j + (long long)j * lda]);
}
__global__ void ainv_kernel(const double *A, long long lda, int p, double *ainv) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < p; j += gridDim.x * blockDim.x)
    ainv[j] = 1.0 / A[j + (long long)j * lda];
}
// max_k |b_k| [/ omega_k]  (_findLambdaMax at x = 0, coordinate_descent.jl:118-149)
__global__ void lambda_max_quad_kernel(const double *b, const double *omega, int p, double *out) {
  __shared__ double red[32];
  double m = 0.0;
  for (int j = threadIdx.x; j < p; j += blockDim.x) {
    double t = fabs(b[j]);
    if (omega) t = t / omega[j];
    if (t > m) m = t;
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
    m = warp_max(m);
    if (threadIdx.x == 0) *out = m;
  }
}
// issymmetric(A) exactly (cd_differentiable_function.jl:306): 32x32 tiles, both reads coalesced
__global__ void symmetric_kernel(const double *A, long long lda, int p, int *flag) {
  __shared__ double t[32][33];
  const int bi = blockIdx.x, bj = blockIdx.y;
  if (bj > bi) return;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int r = ty; r < 32; r += blockDim.y) {
    int i = bj * 32 + tx, j = bi * 32 + r; // tile (bj rows, bi cols)
    t[r][tx] = (i < p && j < p) ? A[i + (long long)j * lda] : 0.0;
  }
  __syncthreads();
  int bad = 0;
  for (int r = ty; r < 32; r += blockDim.y) {
    int i = bi * 32 + tx, j = bj * 32 + r; // tile (bi rows, bj cols): A[i,j] vs A[j,i] = t[tx][r]
    if (i < p && j < p && A[i + (long long)j * lda] != t[tx][r]) bad = 1;
  }
  if (bad) atomicExch(flag, 1);
}

} // namespace

int launch_cov_init(cdgpu_handle_s *h, const double *A, long long lda, int p, const int *act, const double *actval,
                    const int *nact, double *Ax, double *beta, unsigned char *inlist, const int *colslot) {
  int blocks = (p + 255) / 256;
  cov_init_kernel<<<blocks, 256, 0, h->stream>>>(A, lda, p, act, actval, nact, Ax, beta, inlist, colslot);
  scatter_iterate_kernel<<<32, 256, 0, h->stream>>>(act, actval, nact, beta, inlist);
  CUDA_TRY(cudaGetLastError());
  CD_COUNT_LAUNCH(2);
  return CDGPU_OK;
}
int launch_diag_sqrt(cdgpu_handle_s *h, const double *A, long long lda, int p, double *out) {
  diag_sqrt_kernel<<<(p + 255) / 256, 256, 0, h->stream>>>(A, lda, p, out);
  CUDA_TRY(cudaGetLastError());
  CD_COUNT_LAUNCH(1);
  return CDGPU_OK;
}
int launch_extract_ainv(cdgpu_handle_s *h, const double *A, long long lda, int p, double *ainv) {
  ainv_kernel<<<(p + 255) / 256, 256, 0, h->stream>>>(A, lda, p, ainv);
  CUDA_TRY(cudaGetLastError());
  CD_COUNT_LAUNCH(1);
  return CDGPU_OK;
}
int launch_lambda_max_quad(cdgpu_handle_s *h, const double *b, const double *omega, int p, double *out) {
  lambda_max_quad_kernel<<<1, 1024, 0, h->stream>>>(b, omega, p, out);
  CUDA_TRY(cudaGetLastError());
  CD_COUNT_LAUNCH(1);
  return CDGPU_OK;
}
int launch_check_symmetric(cdgpu_handle_s *h, const double *A, long long lda, int p, int *flag) {
  int nb = (p + 31) / 32;
  dim3 grid(nb, nb), block(32, 8);
  symmetric_kernel<<<grid, block, 0, h->stream>>>(A, lda, p, flag);
  CUDA_TRY(cudaGetLastError());
  CD_COUNT_LAUNCH(1);
  return CDGPU_OK;
}

int launch_cov_path(cdgpu_handle_s *h, const CovArgs &a) {
  static bool attr_done[64] = {false}; // function attributes are per device (context), not per process
  const bool known_dev = h->device >= 0 && h->device < 64;
  auto fixed_for = [](int ecap) {
    return (sizeof(Smem) + 15) / 16 * 16 + chain::STAGE_DOUBLES * sizeof(double) +
           (size_t)ecap * (3 * sizeof(double) + sizeof(int) + sizeof(unsigned) + 2 * sizeof(unsigned short));
  };
  const size_t fixed = fixed_for(COV_ACT_CAP);
  const size_t max_dyn = 227 * 1024;
  if (!known_dev || !attr_done[h->device]) {
    CUDA_TRY(cudaFuncSetAttribute(cov_path_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_dyn));
    CUDA_TRY(cudaFuncSetAttribute(cov_path_kernel<false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    CUDA_TRY(cudaFuncSetAttribute(cov_path_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_dyn));
    CUDA_TRY(cudaFuncSetAttribute(cov_path_kernel<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    if (known_dev) attr_done[h->device] = true;
  }
  // largest cluster the device will co-schedule (16 on B200 with the opt-in, else 8)
  static int dev_max_cluster[64] = {0}; // per device: the occupancy query is not free
  if (h->max_cluster == 0 && h->device < 64) h->max_cluster = dev_max_cluster[h->device];
  int C = h->max_cluster;
  if (C == 0) {
    for (int cand : {16, 8, 4, 2, 1}) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cand);
      cfg.blockDim = dim3(COV_T);
      cfg.dynamicSmemBytes = fixed + 1024;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cand;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      int ncl = 0;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, cov_path_kernel<false>, &cfg);
      if (e == cudaSuccess && ncl >= 1) {
        C = cand;
        break;
      }
      (void)cudaGetLastError();
    }
    if (C == 0) return cdgpu_set_error(CDGPU_ECUDA, "no thread-block cluster configuration can be scheduled");
    h->max_cluster = C;
    if (h->device < 64) dev_max_cluster[h->device] = C;
  }
  if (const char *env = getenv("CDGPU_CLUSTER")) {
    int v = atoi(env);
    if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16) C = v < h->max_cluster ? v : h->max_cluster;
  }
  while (C > 1 && a.p < C * 32) C >>= 1; // tiny problems: fewer, fuller slices
  int L = (a.p + C - 1) / C;
  L = (L + 1) & ~1;
  // engine capacity: as large as still leaves room for the slices in shared memory
  // safe screening of the verification sweep: the value is the number of CURRENT rows per CTA that triggers a resync
  // (CDGPU_COV_SCREEN=0: every row verified in every round; 2: diagnostic, every tested row promoted).  C2 sweep kernel:
  // 7.07 / 7.12 / 7.14 / 7.23 ms at 48 / 96 / 160 / 384 against 8.05 ms unscreened in the same build.
  int screen = 64;
  if (const char *env = getenv("CDGPU_COV_SCREEN")) screen = atoi(env);
  if (screen == 1) screen = 64;
  if (L >= 65536 || a.events_only) screen = 0;
  size_t slices = screen ? (size_t)8 * L * sizeof(double) + 5 * (size_t)L + 32 : (size_t)6 * L * sizeof(double) + 2 * (size_t)L + 16;
  int ecap = COV_ACT_CAP;
  if (const char *env = getenv("CDGPU_ECAP_MAX")) ecap = std::max(512, std::min(COV_ACT_CAP, atoi(env) / 512 * 512));
  while (ecap > 1024 && fixed_for(ecap) + slices > max_dyn) ecap >>= 1;
  if (screen && fixed_for(ecap) + slices > max_dyn) { // no room for the screening arrays: plain verification
    screen = 0;
    slices = (size_t)6 * L * sizeof(double) + 2 * (size_t)L + 16;
    ecap = COV_ACT_CAP;
    while (ecap > 1024 && fixed_for(ecap) + slices > max_dyn) ecap >>= 1;
  }
  int slice_in_smem = fixed_for(ecap) + slices <= max_dyn;
  if (!slice_in_smem) {
    ecap = COV_ACT_CAP;
    screen = 0;
  }
  size_t dyn = slice_in_smem ? fixed_for(ecap) + slices : fixed_for(ecap);
  // the distributed engine needs 8p + 80 scratch doubles behind the slices; CDGPU_COV_MULTI=0 switches it off
  int multi_ok = (a.p >= 64 && C >= 2) ? 384 : 0; // run_multi: the chain CTA plus at least one owner CTA
  if (const char *env = getenv("CDGPU_MULTI_MIN")) multi_ok = multi_ok ? std::max(64, atoi(env)) : 0;
  if (const char *env = getenv("CDGPU_COV_MULTI")) multi_ok = atoi(env) != 0 ? multi_ok : 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(C);
  cfg.blockDim = dim3(COV_T);
  cfg.dynamicSmemBytes = dyn;
  cfg.stream = h->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = C;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (a.prof)
    CUDA_TRY(cudaLaunchKernelEx(&cfg, cov_path_kernel<true>, a, L, slice_in_smem, ecap, multi_ok, screen));
  else
    CUDA_TRY(cudaLaunchKernelEx(&cfg, cov_path_kernel<false>, a, L, slice_in_smem, ecap, multi_ok, screen));
  CD_COUNT_LAUNCH(1);
  return CDGPU_OK;
}
