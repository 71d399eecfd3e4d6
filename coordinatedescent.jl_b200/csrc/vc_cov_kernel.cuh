// vc_cov_kernel.cuh — the moment-form kernel of the batched local problems (locpolyl1 / lvocv_locpolyl1), shared by the
// translation units that instantiate it (vc_cov_std.cu, vc_cov_lvo.cu: compiled in parallel) and by the
// host driver in vc_batch.cu.
#pragma once
#include <type_traits>

#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// MOMENT FORM (default for ep <= 256): the covariance form of every local problem at once.
// With eX[i,(j,l)] = X[i,j] dz_i^l the weighted Gram of grid point g is
//     A_g[(j,l),(j',l')] = sum_i w_gi dz_gi^(l+l') X[i,j] X[i,j'] / n = M_{g,l+l'}[j,j'],
// i.e. 2d+1 weighted p x p moment matrices per grid point, and ALL of them for ALL grid points are
// ONE dense FP64 contraction  C = Z'V / n  with  Z[i,(j>=j')] = X[i,j] X[i,j'] | Z[i,P2+j] = X[i,j] y_i
// (n x (p(p+1)/2 + p)) and V[i,(g,q)] = w_gi dz_gi^q (n x m(2d+1)) — run on the FP64 tensor cores by
// the DMMA kernel of gram_dmma.cu (launch_gemm_tn).  The local lasso is then CDQuadraticLoss
// (cd_differentiable_function.jl:324-348) on A_g, b_g = -eX'W y/n with omega_k = sqrt(A_kk)
// (utils.jl:140-151) — the same minimiser as the reference's CDWeightedLSLoss form (:165-194);
// iterates differ at rounding level.  ONE WARP PER GRID POINT: lane owns coordinates lane, lane+32, ...
// with (A x)_t, beta_t and the per-coordinate constants in registers; a full pass is the speculative
// first-mover scan (non-moving coordinates cost no memory traffic), an accepted step reads one
// column of A_g through the moment blocks (L1/L2 resident).
struct VcCovArgs {
  const double *C; // moment blocks: problem g at C + (g - g0) * NQ * ldc, block q at + q * ldc
  long long ldc;
  int p, degree, ep, P2;
  int g0, g1;
  double lambda0;
  long long maxIter;
  double optTol;
  int randomize;
  unsigned long long seed;
  double *out, *outR; // outR: refitted coefficients (null: no refit)
  DevStats *stats;
  int *counter; // dynamic work distribution
  // lvocv_locpolyl1: problem g = (bandwidth g / n, left-out observation g % n)
  int lvo, n;
  const double *X, *y; // device copies (n x p, ldx) for the prediction of the left-out response
  long long ldx;
  double *lvo_err;     // squared prediction error per problem
  unsigned long long *prof; // optional [8]: summed warp cycles: full passes, active chain, list compaction, phase open, phase close, total
  double *gscr; // per-warp scratch for the compact active Gram: (grid * VCW) x MC x MC doubles, MC = ep rounded up to even
  // grid points per work item, >= 1: a warp solves `chain` consecutive grid points one after the other, each from the
  // previous one's iterate (values AND list order: the reference's loop, varying_coefficient_lasso.jl:56,68, is
  // chain = m); 1: every grid point from zero
  int chain;
};

constexpr int VCW = 1;      // warps (local problems) per CTA (one: shared memory then packs 11 problems per SM)
// columns of the compact active Gram in flight to shared memory per warp = template parameter VC_RING of the kernel.
// Measured (round 2, C4): 8 stages instead of 4 change nothing per warp and cost two resident problems per SM
// (31.4 -> 37.9 ms): the chain is bound by its own instruction issue and dependent latencies, not by the column fetch.
__host__ __device__ inline size_t vc_cov_warp_bytes(int ep, int nu, int ring) { // nu: the kernel instance's slots per lane
  const int RS = 32 * nu; // ring stage stride: a whole number of 32-lane rows, so no lane ever clamps
  return ((size_t)(ring * RS + 8 * ep) * sizeof(double) + (size_t)(6 * ep + 4) * sizeof(int) + (size_t)ep + 15) / 16 * 16;
}

// slots per lane of the kernel instance that serves nu = ceil(ep / 32) (VC_DEFINE_PICK below)
__host__ __device__ inline int vc_cov_inst(int nu) { return nu <= 6 ? (nu < 1 ? 1 : nu) : (nu <= 8 ? 8 : (nu <= 12 ? 12 : 16)); }
constexpr int VC_COV_MAX_EP = 512; // widest expanded problem of the moment form (16 slots per lane)

template <int OFF>
__device__ __forceinline__ void vc_cp16(unsigned dst, const char *src) {
  asm volatile("cp.async.cg.shared.global [%0+%2], [%1+%2], 16;" ::"r"(dst), "l"(src), "n"(OFF) : "memory");
}

template <int NU, bool LVO, int VC_RING>
__global__ void __launch_bounds__(VCW * 32, VC_RING > 4 ? 9 : 11) vc_cov_kernel(const VcCovArgs a) {
  extern __shared__ __align__(16) unsigned char raw[];
  const int ep = a.ep, dg = a.degree + 1, nq = 2 * a.degree + 1, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int MC = (ep + 1) & ~1, RS = 32 * NU; // scratch leading dimension bound, ring stage stride
  unsigned char *base = raw + warp * vc_cov_warp_bytes(ep, NU, VC_RING);
  double *ring = reinterpret_cast<double *>(base);       // VC_RING prefetched columns of the compact active Gram
  double *sbeta = ring + VC_RING * RS;                   // dense beta (after a full pass / at phase start)
  double *sval = sbeta + ep, *stmpd = sval + ep;         // list-order values, scratch
  double *sAx = stmpd + ep, *scc = sAx + ep, *sai = scc + ep, *sth = sai + ep; // dense (A x) and constants
  double *sAxE = sth + ep;                               // (A x) of the active entries of a phase, by snapshot entry
  int *sact = reinterpret_cast<int *>(sAxE + ep);
  int *snewpos = sact + ep, *sact0 = snewpos + ep, *stmpi = sact0 + ep, *s2 = stmpi + 2 * ep, *spos = s2 + 4;
  int *scmp = reinterpret_cast<int *>(ring); // cd_compact_list's 5*m ints: the column ring is idle whenever a list is compacted
  unsigned char *sin = reinterpret_cast<unsigned char *>(spos + ep);
  double *Gw = a.gscr + ((long long)blockIdx.x * VCW + warp) * (long long)MC * MC; // this warp's compact Gram scratch
  const bool ordered = a.randomize == 0;
  constexpr unsigned NONE = 0xffffffffu;

  int tj[NU], tl[NU], ttri[NU];
#pragma unroll
  for (int u = 0; u < NU; ++u) {
    const int t = lane + 32 * u;
    tj[u] = t / dg;
    tl[u] = t - tj[u] * dg;
    ttri[u] = tj[u] * (tj[u] + 1) / 2;
  }

  double be[NU]; // the iterate in coordinate layout: carried from one grid point of a chain to the next
  int nact = 0;
  for (;;) {
    int wi = 0;
    if (lane == 0) wi = atomicAdd(a.counter, 1);
    wi = __shfl_sync(0xffffffffu, wi, 0);
    const long long gfirst = (long long)a.g0 + (long long)wi * a.chain;
    if (gfirst >= a.g1) break;
    const int glast = (int)min((long long)a.g1, gfirst + a.chain);
    for (int g = (int)gfirst; g < glast; ++g) {
    const bool cold = g == (int)gfirst;
    const double *Cg = a.C + (long long)(g - a.g0) * nq * a.ldc;
    // element (k1, k2) of A_g through the moment blocks
    auto moment = [&](int k1, int k2) -> double {
      const int j1 = k1 / dg, l1 = k1 - j1 * dg, j2 = k2 / dg, l2 = k2 - j2 * dg;
      const int pk_ = j1 >= j2 ? j1 * (j1 + 1) / 2 + j2 : j2 * (j2 + 1) / 2 + j1;
      return __ldg(Cg + (long long)(l1 + l2) * a.ldc + pk_);
    };
    // state: COORDINATE layout outside a phase (slot u <-> coordinate lane + 32u), ENTRY layout inside one
    // (slot u <-> snapshot entry lane + 32u of the phase's active list)
    double Ax[NU], cc[NU], ai[NU], th[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      const int t = lane + 32 * u;
      Ax[u] = cc[u] = th[u] = 0.0;
      if (cold) be[u] = 0.0;
      ai[u] = 0.0;
      if (t < ep) {
        const double att = __ldg(Cg + (long long)(2 * tl[u]) * a.ldc + ttri[u] + tj[u]);
        ai[u] = 1.0 / att;
        th[u] = __dmul_rn(__dmul_rn(ai[u], a.lambda0), sqrt(att)); // lambda0 * omega_k / A_kk, omega_k = sqrt(A_kk)
        cc[u] = -__ldg(Cg + (long long)tl[u] * a.ldc + a.P2 + tj[u]);
        scc[t] = cc[u];
        sai[t] = ai[u];
        sth[t] = th[u];
      }
    }
    if (cold) {
      for (int k = lane; k < ep; k += 32) sin[k] = 0;
      nact = 0;
    }
    __syncwarp();

    // coordinate layout: Ax[u] += A_g[t_u, k] * h for the mover k = (kj, kl)
    auto apply = [&](int k, double h) {
      const int kj = k / dg, kl = k - kj * dg, ktri = kj * (kj + 1) / 2;
      double gv[NU];
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        const int t = lane + 32 * u;
        const int pk_ = kj >= tj[u] ? ktri + tj[u] : ttri[u] + kj;
        gv[u] = t < ep ? __ldg(Cg + (long long)(kl + tl[u]) * a.ldc + pk_) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < NU; ++u) Ax[u] = __dadd_rn(Ax[u], __dmul_rn(gv[u], h));
    };

    // warm start inside a chain: (A x) of this grid point's Gram at the carried iterate, stored entries in list order
    // (what initialize! does for the reference's loss, cd_differentiable_function.jl:135-148)
    if (!cold)
      for (int i = 0; i < nact; ++i) apply(sact[i], sval[i]);

    // ---- phases of consecutive active-set passes.  Active-set passes only need (A x) on the active set, so a
    // phase (a) gathers the compact m0 x m0 block A_g[act0, act0] once into this warp's scratch (L2 resident,
    // column-contiguous), (b) moves the state of the active entries into the entry layout, (c) runs the chain
    // with the column of the next VC_RING-1 steps already in flight to shared memory (cp.async), so a step is:
    // owner evaluates, one shuffle broadcasts h, every lane updates its entries — no memory latency on the
    // dependent path, and (d) at the end brings (A x) of the other coordinates up to date from the change of beta.
    bool in_phase = false;
    int m0 = 0, ldw = 0;
    auto phase_open = [&](int m) {
      m0 = m;
      ldw = (m + 1) & ~1;
#pragma unroll
      for (int u = 0; u < NU; ++u)
        if (lane + 32 * u < ep) {
          sbeta[lane + 32 * u] = be[u]; // beta at the start of the phase
          sAx[lane + 32 * u] = Ax[u];   // (A x) at the start of the phase, by coordinate
        }
      for (int k = lane; k < ep; k += 32) spos[k] = -1;
      __syncwarp();
      for (int i = lane; i < m; i += 32) {
        const int k = sact[i];
        sact0[i] = k;
        spos[k] = i;
        sAxE[i] = sAx[k]; // (A x) of the active entries, by snapshot entry, carried from pass to pass
      }
      __syncwarp();
      // compact Gram: 4 columns per iteration so that 4 * ceil(m/32) independent gathers are in flight per lane
      for (int j = 0; j < m; j += 4) {
        int kj[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) kj[c] = sact0[min(j + c, m - 1)];
        for (int i = lane; i < m; i += 32) {
          const int ki = sact0[i];
          double v[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) v[c] = moment(ki, kj[c]);
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (j + c < m) Gw[i + (long long)(j + c) * ldw] = v[c];
        }
      }
      __syncwarp();
      in_phase = true;
    };
    // back to the coordinate layout.  Coordinates in the list take their tracked (A x); every other coordinate
    // (never active in this phase, or dropped during it) gets (A x) at phase start + sum_e A[t, act0_e] (beta_e -
    // beta_e at phase start).
    auto phase_close = [&](int nact_now) {
      for (int k = lane; k < ep; k += 32) stmpd[k] = 0.0; // dense beta now (dropped entries are exactly zero)
      __syncwarp();
      for (int i = lane; i < nact_now; i += 32) {
        const int k = sact[i];
        stmpd[k] = sval[i];
        sAx[k] = sAxE[spos[k]];
      }
      __syncwarp();
      bool inactive[NU];
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        const int t = lane + 32 * u;
        Ax[u] = be[u] = cc[u] = ai[u] = th[u] = 0.0;
        inactive[u] = false;
        if (t < ep) {
          inactive[u] = sin[t] == 0;
          Ax[u] = sAx[t];
          be[u] = stmpd[t];
          cc[u] = scc[t];
          ai[u] = sai[t];
          th[u] = sth[t];
        }
      }
      for (int e0 = 0; e0 < m0; e0 += 2) { // two entries per iteration: twice the loads in flight
        double gv[2][NU], dl[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int k = sact0[min(e0 + c, m0 - 1)];
          dl[c] = e0 + c < m0 ? stmpd[k] - sbeta[k] : 0.0;
          const int kj = k / dg, kl = k - kj * dg, ktri = kj * (kj + 1) / 2;
#pragma unroll
          for (int u = 0; u < NU; ++u) {
            const int pk_ = kj >= tj[u] ? ktri + tj[u] : ttri[u] + kj;
            gv[c][u] = inactive[u] && dl[c] != 0.0 ? __ldg(Cg + (long long)(kl + tl[u]) * a.ldc + pk_) : 0.0;
          }
        }
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int u = 0; u < NU; ++u)
            if (inactive[u]) Ax[u] = fma(gv[c][u], dl[c], Ax[u]);
      }
      __syncwarp();
      in_phase = false;
    };
    // One active-set pass inside a phase; returns max|h|.  The m listed entries are laid out IN VISIT ORDER for the
    // pass: visit position s <-> (lane s & 31, slot s >> 5), so the stepping slot is a compile-time index and the
    // owner is the loop counter: no selects, no uniform loads on the chain.  A step: every lane evaluates its own
    // slot-u entry, lane i's h is broadcast by one shuffle, every lane updates all its entries with the column of
    // the compact Gram that the cp.async ring brought to shared memory VC_RING-1 steps earlier (rows are read
    // through the lane's position -> snapshot-entry map).
    // MS = ceil(m / 32) slots per lane are live in the pass: the per-step work (gathers, multiply-adds) is instantiated
    // for exactly that many, instead of the NU the widest possible list would need
    auto phase_pass = [&](int m, const PermKey &pkm, int &accepted, auto ms_tag) -> double {
      constexpr int MS = decltype(ms_tag)::value;
      unsigned *soff = reinterpret_cast<unsigned *>(stmpi); // byte offset of the column of visit position s in Gw
      int *slist = stmpi + ep;                              // list position i_ of visit position s
      for (int s_ = lane; s_ < m; s_ += 32) {
        const int i_ = ordered ? s_ : (int)cd_perm(pkm, (uint32_t)s_);
        slist[s_] = i_;
        soff[s_] = (unsigned)(spos[sact[i_]] * ldw) * 8u;
      }
      __syncwarp();
      int row[MS]; // snapshot entry (= row of the compact Gram) of this lane's visit positions
#pragma unroll
      for (int u = 0; u < MS; ++u) {
        const int s_ = lane + 32 * u;
        Ax[u] = be[u] = cc[u] = ai[u] = th[u] = 0.0;
        row[u] = 0;
        if (s_ < m) {
          const int i_ = slist[s_], k = sact[i_];
          row[u] = spos[k];
          Ax[u] = sAxE[row[u]];
          be[u] = sval[i_];
          cc[u] = scc[k];
          ai[u] = sai[k];
          th[u] = sth[k];
        }
      }
      const char *gsrc = reinterpret_cast<const char *>(Gw) + lane * 16;
      const unsigned rdst = (unsigned)__cvta_generic_to_shared(ring) + lane * 16;
      auto fetch = [&](int s_) { // column of visit position s_ -> ring stage s_ % VC_RING (16-byte chunks)
        if (s_ < m) {
          const char *src = gsrc + soff[s_];
          const unsigned dst = rdst + (unsigned)(s_ % VC_RING) * (unsigned)(RS * 8);
          if (2 * lane < m0) vc_cp16<0>(dst, src);
          if ((NU + 1) / 2 > 1 && 2 * (lane + 32) < m0) vc_cp16<512>(dst, src);
          if ((NU + 1) / 2 > 2 && 2 * (lane + 64) < m0) vc_cp16<1024>(dst, src);
          if ((NU + 1) / 2 > 3 && 2 * (lane + 96) < m0) vc_cp16<1536>(dst, src);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      };
#pragma unroll
      for (int d = 0; d < VC_RING - 1; ++d) fetch(d);
      double maxH = 0.0;
#pragma unroll
      for (int u = 0; u < MS; ++u) {
        const int cnt = min(32, m - 32 * u);
        for (int i = 0; i < cnt; ++i) { // visit position s_ = 32 u + i, ring stage i % VC_RING (32 % VC_RING == 0)
          const int s_ = 32 * u + i;
          asm volatile("cp.async.wait_group %0;" ::"n"(VC_RING - 2) : "memory"); // column of position s_ has landed
          __syncwarp();
          const double *col = ring + (i % VC_RING) * RS;
          double gc[MS];
#pragma unroll
          for (int q = 0; q < MS; ++q) gc[q] = col[row[q]];
          fetch(s_ + VC_RING - 1); // into the stage of position s_-1, which every lane read before the barrier above
          const double v = __dsub_rn(be[u], __dmul_rn(Ax[u] + cc[u], ai[u]));
          const double nwl = cd_shrink(v, th[u]);
          const double h = __shfl_sync(0xffffffffu, nwl - be[u], i);
          if (lane == i) be[u] = nwl;
          // h == 0 adds an exact zero (the Gram entries are finite): no data-dependent branch on the chain
#pragma unroll
          for (int q = 0; q < MS; ++q) Ax[q] = __dadd_rn(Ax[q], __dmul_rn(gc[q], h));
          accepted += h != 0.0;
          maxH = fmax(maxH, fabs(h));
        }
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
#pragma unroll
      for (int u = 0; u < MS; ++u) {
        const int s_ = lane + 32 * u;
        if (s_ < m) {
          sAxE[row[u]] = Ax[u];
          sval[slist[s_]] = be[u];
        }
      }
      __syncwarp();
      return maxH;
    };

    DevStats st;
    st.passes = st.full_passes = st.visits = st.accepted = 0;
    st.maxH = 0.0;
    st.converged = 0;
    st.outer_iters = 0;
    st.sigma = 0.0;
    unsigned long long pass_counter = 0;
    long long pc[6] = {0, 0, 0, 0, 0, 0};
    const long long tstart = clock64();
    // _coordinateDescent! (coordinate_descent.jl:65-92) from the current iterate: full pass first, active-set passes
    // until one converges, exit when a full pass has max|h| < optTol
    auto run_solve = [&]() {
    bool conv = true;
    long long iter = 0;
    st.converged = 0;
    while (iter < a.maxIter) {
      double maxH = 0.0;
      iter += 1;
      st.passes += 1;
      long long tq = clock64();
      if (conv) { // ---- full pass: speculative first-mover scan (exact Gauss-Seidel order)
        if (in_phase) {
          phase_close(nact);
          pc[4] += clock64() - tq;
          tq = clock64();
        }
        st.full_passes += 1;
        st.visits += ep;
        const PermKey pk = cd_perm_key((uint32_t)ep, a.seed, pass_counter);
        const int m_old = nact;
        long long cur = -1;
        for (;;) {
          unsigned best = NONE;
          double bh = 0.0, bnw = 0.0;
          int bk = 0;
#pragma unroll
          for (int u = 0; u < NU; ++u) {
            const int t = lane + 32 * u;
            if (t < ep) {
              const unsigned key = ordered ? (unsigned)t : cd_perm_inv(pk, (uint32_t)t);
              if ((long long)key > cur) {
                const double tt = __dmul_rn(Ax[u] + cc[u], ai[u]);
                const double v = __dsub_rn(be[u], tt);
                const double nw = cd_shrink(v, th[u]);
                const double h = nw - be[u];
                if (h != 0.0 && key < best) {
                  best = key;
                  bh = h;
                  bnw = nw;
                  bk = t;
                }
              }
            }
          }
          const unsigned wmin = __reduce_min_sync(0xffffffffu, best);
          if (wmin == NONE) break;
          const int src = __ffs(__ballot_sync(0xffffffffu, best == wmin)) - 1;
          const int k = __shfl_sync(0xffffffffu, bk, src);
          const double h = __shfl_sync(0xffffffffu, bh, src), nw = __shfl_sync(0xffffffffu, bnw, src);
          if (lane == (k & 31)) {
#pragma unroll
            for (int u = 0; u < NU; ++u)
              if (u == (k >> 5)) be[u] = nw;
          }
          apply(k, h);
          if (!sin[k]) { // setindex! appends on the first non-zero store
            __syncwarp();
            if (lane == 0) {
              sin[k] = 1;
              sact[nact] = k;
            }
            nact += 1;
            __syncwarp();
          }
          maxH = fmax(maxH, fabs(h));
          st.accepted += 1;
          cur = (long long)wmin;
        }
        pc[0] += clock64() - tq;
        tq = clock64();
        // list order after the pass (common.cuh: cd_compact_list); the visited non-members whose tentative
        // value is exactly zero (not appended by the reference) are not tracked here: it needs an exactly
        // zero gradient and only affects the visit order of later active-set passes
#pragma unroll
        for (int u = 0; u < NU; ++u)
          if (lane + 32 * u < ep) sbeta[lane + 32 * u] = be[u];
        __syncwarp();
        for (int i = lane; i < nact; i += 32) sval[i] = sbeta[sact[i]];
        for (int e = m_old + lane; e < nact; e += 32) {
          const int k = sact[e];
          const int vis = ordered ? k : (int)cd_perm_inv(pk, (uint32_t)k);
          int before = 0;
          for (int j = 0; j < m_old; ++j) before += (ordered ? sact[j] : (int)cd_perm_inv(pk, (uint32_t)sact[j])) < vis;
          snewpos[e - m_old] = m_old + vis - before;
        }
        __syncwarp();
        cd_compact_list<32>(sact, sval, m_old, nact, snewpos, sin, scmp, stmpd, s2);
        nact = s2[0];
        __syncwarp();
        pc[2] += clock64() - tq;
      } else { // ---- active-set pass: sequential chain over the stored entries
        const int m = nact;
        st.visits += m;
        if (!in_phase) {
          phase_open(m);
          pc[3] += clock64() - tq;
          tq = clock64();
        }
        const PermKey pkm = cd_perm_key((uint32_t)max(m, 1), a.seed, pass_counter);
        int acc_pass = 0;
        {
          const int ms = (m + 31) >> 5;
          if (ms <= 1 || NU == 1)
            maxH = phase_pass(m, pkm, acc_pass, std::integral_constant<int, 1>{});
          else if (ms == 2 || NU == 2)
            maxH = phase_pass(m, pkm, acc_pass, std::integral_constant<int, (NU < 2 ? NU : 2)>{});
          else if (ms == 3 || NU == 3)
            maxH = phase_pass(m, pkm, acc_pass, std::integral_constant<int, (NU < 3 ? NU : 3)>{});
          else if (ms == 4 || NU == 4)
            maxH = phase_pass(m, pkm, acc_pass, std::integral_constant<int, (NU < 4 ? NU : 4)>{});
          else
            maxH = phase_pass(m, pkm, acc_pass, std::integral_constant<int, NU>{});
        }
        st.accepted += acc_pass;
        pc[1] += clock64() - tq;
        tq = clock64();
        cd_compact_list<32>(sact, sval, m, m, snewpos, sin, scmp, stmpd, s2); // dropzeros!
        nact = s2[0];
        __syncwarp();
        pc[2] += clock64() - tq;
      }
      pass_counter += 1;
      st.maxH = maxH;
      const bool prev = conv;
      conv = maxH < a.optTol;
      if (prev && conv) {
        st.converged = 1;
        break;
      }
    }
    if (in_phase) phase_close(nact); // pass budget ran out inside a phase: back to the coordinate layout
    };

    // A_g[S,S] x = rhs for the ms coordinates listed (ascending) in sact; rhs and the solution live in stmpd.
    // Left-looking Cholesky in this warp's scratch (columns contiguous: coalesced; all loads of a column's update
    // are independent), then the two triangular solves.  false: not positive definite.
    auto spd_solve = [&](int ms) -> bool {
      const int ldm = (ms + 1) & ~1;
      double *rhs = stmpd, *lrow = sval; // lrow: L[j, 0..j) of the column being formed
      bool ok = true;
      for (int j = 0; j < ms; ++j) {
        const int kj = sact[j];
        for (int k2 = lane; k2 < j; k2 += 32) lrow[k2] = Gw[j + (long long)k2 * ldm];
        __syncwarp();
        double acc[NU];
#pragma unroll
        for (int u = 0; u < NU; ++u) {
          const int i = j + lane + 32 * u;
          acc[u] = i < ms ? moment(sact[i], kj) : 0.0;
        }
        for (int k2 = 0; k2 < j; ++k2) {
          const double ljk = lrow[k2];
          const double *ck = Gw + (long long)k2 * ldm + j + lane;
#pragma unroll
          for (int u = 0; u < NU; ++u)
            if (j + lane + 32 * u < ms) acc[u] = fma(-__ldcg(ck + 32 * u), ljk, acc[u]);
        }
        const double djj = __shfl_sync(0xffffffffu, acc[0], 0);
        if (!(djj > 0.0)) ok = false;
        const double d = sqrt(djj);
#pragma unroll
        for (int u = 0; u < NU; ++u) {
          const int i = j + lane + 32 * u;
          if (i < ms) Gw[i + (long long)j * ldm] = i == j ? d : acc[u] / d;
        }
        __syncwarp();
      }
      for (int j = 0; j < ms; ++j) { // L y = rhs (column sweep)
        const double yj = rhs[j] / __ldcg(Gw + j + (long long)j * ldm);
        __syncwarp();
        if (lane == 0) rhs[j] = yj;
        for (int i = j + 1 + lane; i < ms; i += 32) rhs[i] = fma(-__ldcg(Gw + i + (long long)j * ldm), yj, rhs[i]);
        __syncwarp();
      }
      for (int j = ms - 1; j >= 0; --j) { // L' x = y (dot products)
        double sacc = 0.0;
        for (int i = j + 1 + lane; i < ms; i += 32) sacc = fma(__ldcg(Gw + i + (long long)j * ldm), rhs[i], sacc);
        sacc = warp_sum(sacc);
        const double xj = (rhs[j] - sacc) / __ldcg(Gw + j + (long long)j * ldm);
        __syncwarp();
        if (lane == 0) rhs[j] = xj;
        __syncwarp();
      }
      return ok;
    };
    // the expanded coordinates of every group with a non-zero coefficient, ascending, into sact
    // (get_nonzero_coordinates!(S, beta, p, degree, true), varying_coefficient_lasso.jl:488-512); returns |S|
    auto selected_groups = [&]() -> int {
      for (int k2 = lane; k2 < ep; k2 += 32) sin[k2] = 0;
      __syncwarp();
#pragma unroll
      for (int u = 0; u < NU; ++u)
        if (lane + 32 * u < ep && be[u] != 0.0) {
          const int j0 = tj[u] * dg;
          for (int l = 0; l < dg; ++l) sin[j0 + l] = 1; // the whole group (same byte value from every writer)
        }
      __syncwarp();
      int ms = 0;
      for (int k0 = 0; k0 < ep; k0 += 32) {
        const int k2 = k0 + lane;
        const bool inS = k2 < ep && sin[k2] != 0;
        const unsigned bal = __ballot_sync(0xffffffffu, inS);
        if (inS) sact[ms + __popc(bal & ((1u << lane) - 1u))] = k2;
        ms += __popc(bal);
      }
      __syncwarp();
      return ms;
    };

    if constexpr (!LVO) {
      run_solve();
    } else {
      // ---- one problem of lvocv_locpolyl1 (varying_coefficient_lasso.jl:81-137): scaled-lasso sigma loop on the
      // leave-one-out local problem, refit, prediction of the left-out response.  Everything about the residual
      // comes from the moment blocks: r'Wr/n = y'Wy/n + 2 b'beta + beta'(A beta), sum(w)/n.
      const double yWy = __ldg(Cg + a.P2 + a.p), Sw = __ldg(Cg + a.P2 + a.p + 1);
      auto sigma_now = [&]() -> double { // _getSigma(w, f.r), utils.jl:167-175
        double t = 0.0;
#pragma unroll
        for (int u = 0; u < NU; ++u) t += be[u] * (2.0 * cc[u] + Ax[u]);
        t = warp_sum(t);
        return sqrt(fmax(yWy + t, 0.0) / Sw);
      };
      // _findInitResiduals!(w, wX, y, min(10, ep), f.r) (utils.jl:79-92): the s columns with the largest |X_k'Wy|,
      // weighted least squares on them, sigma from those residuals
      const int s_init = min(10, ep);
      double thr = 0.0;
      {
        double cv[NU];
#pragma unroll
        for (int u = 0; u < NU; ++u) cv[u] = lane + 32 * u < ep ? fabs(cc[u]) : -1.0;
        for (int r = 0; r < s_init; ++r) { // r-th largest by repeated arg-max
          double best = -1.0;
#pragma unroll
          for (int u = 0; u < NU; ++u) best = fmax(best, cv[u]);
          const double wbest = warp_max(best);
          thr = wbest;
          const unsigned bal = __ballot_sync(0xffffffffu, best == wbest);
          if (lane == __ffs(bal) - 1) {
            bool done = false;
#pragma unroll
            for (int u = 0; u < NU; ++u)
              if (!done && cv[u] == wbest) {
                cv[u] = -1.0;
                done = true;
              }
          }
        }
      }
      int ms0 = 0;
      for (int k0 = 0; k0 < ep; k0 += 32) { // S = storage .>= nlargest(s, storage)[end] (ties included)
        const int k2 = k0 + lane;
        const bool inS = k2 < ep && fabs(scc[k2]) >= thr;
        const unsigned bal = __ballot_sync(0xffffffffu, inS);
        if (inS) sact[ms0 + __popc(bal & ((1u << lane) - 1u))] = k2;
        ms0 += __popc(bal);
      }
      __syncwarp();
      for (int i = lane; i < ms0; i += 32) stmpd[i] = -scc[sact[i]];
      __syncwarp();
      spd_solve(ms0);
      double t0 = 0.0;
      for (int i = lane; i < ms0; i += 32) t0 = fma(scc[sact[i]], stmpd[i], t0); // at the LS solution r'Wr/n = y'Wy/n + b_S'gamma
      t0 = warp_sum(t0);
      double sigma = sqrt(fmax(yWy + t0, 0.0) / Sw);
      __syncwarp();
      for (int outer = 1; outer <= 10; ++outer) { // :115-124
        const double lam = a.lambda0 * sigma;
#pragma unroll
        for (int u = 0; u < NU; ++u)
          if (lane + 32 * u < ep) {
            th[u] = __dmul_rn(__dmul_rn(ai[u], lam), sqrt(1.0 / ai[u])); // lambda0 sigma omega_k / A_kk
            sth[lane + 32 * u] = th[u];
          }
        __syncwarp();
        run_solve();
        st.outer_iters = outer;
        const double snew = sigma_now();
        if (fabs(snew - sigma) / sigma < 1e-2) break;
        sigma = snew;
      }
      st.sigma = sigma;
    }
    if (a.prof && lane == 0) {
      pc[5] = clock64() - tstart;
      for (int i = 0; i < 6; ++i) atomicAdd(a.prof + i, (unsigned long long)pc[i]);
      atomicMax(a.prof + 6, (unsigned long long)pc[5]);
    }
    if (a.out) {
      double *col = a.out + (long long)g * ep;
#pragma unroll
      for (int u = 0; u < NU; ++u)
        if (lane + 32 * u < ep) col[lane + 32 * u] = be[u];
    }
    if (lane == 0 && a.stats) a.stats[g] = st;
    __syncwarp();
    const bool keep_list = g + 1 < glast; // the next grid point of the chain starts from this list
    if (LVO || a.outR) {
      if (keep_list) { // the refit reuses sact / sin / sval
        for (int i = lane; i < nact; i += 32) stmpi[i] = sact[i];
        __syncwarp();
      }
      // ---- refit (varying_coefficient_lasso.jl:71-76): A_g[S,S] x = -b_g[S] on the expanded coordinates S of every
      // group with a non-zero coefficient; both sides come from the moment blocks
      const int ms = selected_groups();
      for (int i = lane; i < ms; i += 32) stmpd[i] = -scc[sact[i]];
      __syncwarp();
      const bool ok = ms > 0 ? spd_solve(ms) : true;
      if (a.outR) {
        double *colR = a.outR + (long long)g * ep;
        for (int k2 = lane; k2 < ep; k2 += 32) colR[k2] = 0.0;
        __syncwarp();
        for (int i = lane; i < ms; i += 32) colR[sact[i]] = ok ? stmpd[i] : nan("");
      }
      if constexpr (LVO) {
        // prediction of the left-out response (:129-131): z0 = z_i, so only the degree-0 columns of row i are
        // non-zero: Yh = sum_{(j,0) in S} X[i,j] x_(j,0)
        const int i_obs = g % a.n;
        double yh = 0.0;
        for (int i = lane; i < ms; i += 32) {
          const int k2 = sact[i], j = k2 / dg;
          if (k2 - j * dg == 0) yh = fma(__ldg(a.X + i_obs + (long long)j * a.ldx), stmpd[i], yh);
        }
        yh = warp_sum(yh);
        if (lane == 0) {
          const double e = yh - __ldg(a.y + i_obs);
          a.lvo_err[g] = ok ? e * e : nan("");
        }
      }
      __syncwarp();
      if (keep_list) {
        for (int k2 = lane; k2 < ep; k2 += 32) sin[k2] = 0;
#pragma unroll
        for (int u = 0; u < NU; ++u)
          if (lane + 32 * u < ep) sbeta[lane + 32 * u] = be[u];
        __syncwarp();
        for (int i = lane; i < nact; i += 32) {
          const int k2 = stmpi[i];
          sact[i] = k2;
          sin[k2] = 1;
          sval[i] = sbeta[k2];
        }
        __syncwarp();
      }
    }
    } // grid points of the chain
  }
}


} // namespace

// one translation unit per (LVO, RING) pair: const void *NAME(int nu) returns the kernel instance for nu slots per lane
#define VC_DEFINE_PICK(NAME, LV, RG)                                       \
  const void *NAME(int nu) {                                               \
    return nu <= 1   ? (const void *)vc_cov_kernel<1, LV, RG>              \
           : nu == 2 ? (const void *)vc_cov_kernel<2, LV, RG>              \
           : nu == 3 ? (const void *)vc_cov_kernel<3, LV, RG>              \
           : nu == 4 ? (const void *)vc_cov_kernel<4, LV, RG>              \
           : nu == 5 ? (const void *)vc_cov_kernel<5, LV, RG>              \
           : nu == 6 ? (const void *)vc_cov_kernel<6, LV, RG>              \
           : nu <= 8 ? (const void *)vc_cov_kernel<8, LV, RG>              \
           : nu <= 12 ? (const void *)vc_cov_kernel<12, LV, RG>            \
                      : (const void *)vc_cov_kernel<16, LV, RG>;           \
  }
