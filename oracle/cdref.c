/*
 * cdref.c — CPU ORACLE for the coordinate-descent hot path of CoordinateDescent.jl.
 *
 * THIS FILE IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (libcdgpu.so) never links, loads or calls anything in oracle/.
 *
 * It is a plain-C restatement of the reference's Julia algorithm, loop for loop
 * (including the reference's unconditional O(n)/O(p) axpy when h == 0 and its
 * recomputation of the column norm on every visit), exported behind the same C
 * ABI as include/cdgpu.h with the prefix cdref_.  Each function cites the
 * reference file:line it follows (paths relative to /root/reference/).
 *
 * Julia itself is not available in the build container and the reference holds
 * no stored golden vectors (every random test draws from Julia's RNG), so the
 * oracle is pinned by (tests/test_oracle_*.py):
 *   - the reference's known-answer tests: test/coordinate_descent.jl:13-25,
 *     test/atom_iterator.jl:11-48, test/varying_coefficient_lasso.jl:16-21,43-66;
 *   - the reference's solver-independent properties: KKT equalities, LS ==
 *     covariance form, scalar == ones-weighted, path == pointwise, warm == cold ==
 *     randomized (test/lasso.jl, test/coordinate_descent.jl);
 *   - scikit-learn's Lasso(fit_intercept=False), whose objective is exactly
 *     CDLeastSquaresLoss + ProxL1.
 *
 * Third-party arithmetic NOT under /root/reference: ProximalBase v0.3.0
 * (Manifest.toml:68-72, git-tree-sha1 de80bb57c611a91a38f99708f4cbee276c17743d).
 * Its published behaviour is restated in the "ProximalBase" section below:
 *   shrink(v,c) = v>c ? v-c : (v<-c ? v+c : 0);
 *   cdprox!(ProxL1(l0), x,k,g)     : x[k] = shrink(x[k], g*l0)
 *   cdprox!(ProxL1(l0,l), x,k,g)   : x[k] = shrink(x[k], g*l0*l[k])
 *   SparseIterate: dense-capacity nzval/nzval2ind/ind2nzval + nnz; setindex!
 *   appends on the first non-zero store and keeps explicit zeros for present
 *   keys; dropzeros! compacts by moving the last stored entry into the hole.
 * The insertion order is pinned by test/atom_iterator.jl:13-28; the compaction
 * order of dropzeros! and shrink's NaN behaviour are parity-UNPINNED (they are a
 * recollection of the package source, not verifiable offline).  Solutions at
 * convergence do not depend on either.
 *
 * Build: see oracle/Makefile.  libcdref.so      = -O2 -ffp-contract=off (parity)
 *                              libcdref_fast.so = -O3 -march=x86-64-v3   (timing)
 */
#include "../include/cdgpu.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define API __attribute__((visibility("default")))

static __thread char g_err[512];
static int fail(int code, const char *msg) {
  snprintf(g_err, sizeof g_err, "%s", msg);
  return code;
}

static double now_ms(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

/* ===================================================================== */
/* ProximalBase v0.3.0 (external; behaviour restated, see header)         */
/* ===================================================================== */
typedef struct {
  int64_t p;
  double *nzval;      /* 0-based storage of Julia's nzval[1:nnz]          */
  int64_t *nzval2ind; /* 1-based coordinate index of stored entry          */
  int64_t *ind2nzval; /* [k-1] -> 1-based slot, 0 == absent                */
  int64_t nnz;
} sparse_iterate;

static int si_alloc(sparse_iterate *x, int64_t p) {
  x->p = p;
  x->nnz = 0;
  x->nzval = (double *)calloc((size_t)p, sizeof(double));
  x->nzval2ind = (int64_t *)calloc((size_t)p, sizeof(int64_t));
  x->ind2nzval = (int64_t *)calloc((size_t)p, sizeof(int64_t));
  return (x->nzval && x->nzval2ind && x->ind2nzval) ? 0 : -1;
}
static void si_free(sparse_iterate *x) {
  free(x->nzval);
  free(x->nzval2ind);
  free(x->ind2nzval);
}
/* getindex: 0 when absent */
static inline double si_get(const sparse_iterate *x, int64_t k) {
  int64_t s = x->ind2nzval[k - 1];
  return s ? x->nzval[s - 1] : 0.0;
}
/* setindex!: append on the first non-zero store, overwrite (zeros too) when present */
static inline void si_set(sparse_iterate *x, int64_t k, double v) {
  int64_t s = x->ind2nzval[k - 1];
  if (s == 0) {
    if (v != 0.0) {
      x->nnz += 1;
      x->nzval[x->nnz - 1] = v;
      x->nzval2ind[x->nnz - 1] = k;
      x->ind2nzval[k - 1] = x->nnz;
    }
  } else {
    x->nzval[s - 1] = v;
  }
}
/* dropzeros!: the last stored entry moves into the hole (parity-unpinned order) */
static void si_dropzeros(sparse_iterate *x) {
  int64_t i = 1;
  while (i <= x->nnz) {
    if (x->nzval[i - 1] == 0.0) {
      x->ind2nzval[x->nzval2ind[i - 1] - 1] = 0;
      if (i != x->nnz) {
        x->nzval[i - 1] = x->nzval[x->nnz - 1];
        x->nzval2ind[i - 1] = x->nzval2ind[x->nnz - 1];
        x->ind2nzval[x->nzval2ind[i - 1] - 1] = i;
      }
      x->nnz -= 1;
    } else {
      i += 1;
    }
  }
}
static void si_fill_zero(sparse_iterate *x) { /* fill!(x, 0) */
  for (int64_t i = 0; i < x->nnz; ++i) x->ind2nzval[x->nzval2ind[i] - 1] = 0;
  x->nnz = 0;
}
/* load the caller's triple; out-of-range or duplicate keys -> error */
static int si_load(sparse_iterate *x, const double *nzval, const int64_t *nzval2ind, int64_t nnz) {
  if (nnz < 0 || nnz > x->p) return -1;
  si_fill_zero(x);
  for (int64_t i = 0; i < nnz; ++i) {
    int64_t k = nzval2ind[i];
    if (k < 1 || k > x->p || x->ind2nzval[k - 1] != 0) return -1;
    x->nzval[i] = nzval[i];
    x->nzval2ind[i] = k;
    x->ind2nzval[k - 1] = i + 1;
  }
  x->nnz = nnz;
  return 0;
}
static void si_store(const sparse_iterate *x, double *nzval, int64_t *nzval2ind, int64_t *nnz) {
  for (int64_t i = 0; i < x->nnz; ++i) {
    nzval[i] = x->nzval[i];
    nzval2ind[i] = x->nzval2ind[i];
  }
  *nnz = x->nnz;
}

typedef struct {
  double l0;
  const double *l; /* NULL == ProxL1{T,Nothing} */
} prox_l1;

static inline double shrink(double v, double c) { return v > c ? v - c : (v < -c ? v + c : 0.0); }
/* cdprox!(g, x, k, gamma): returns the new x[k] */
static inline double cdprox(const prox_l1 *g, sparse_iterate *x, int64_t k, double gamma) {
  double c = gamma * g->l0;
  if (g->l) c = c * g->l[k - 1];
  double v = shrink(si_get(x, k), c);
  si_set(x, k, v);
  return v;
}
/* A_mul_B_row(X, x, i) = sum_{k in nz} X[i,k] x[k], in storage order */
static inline double A_mul_B_row(const double *X, int64_t ld, const sparse_iterate *x, int64_t i) {
  double v = 0.0;
  for (int64_t s = 0; s < x->nnz; ++s) v += X[(i - 1) + (x->nzval2ind[s] - 1) * ld] * x->nzval[s];
  return v;
}
/* At_mul_B_row(X, r, j) = X[:,j]' r */
static inline double At_mul_B_row(const double *X, int64_t ld, int64_t n, const double *r, int64_t j) {
  double v = 0.0;
  const double *c = X + (j - 1) * ld;
  for (int64_t i = 0; i < n; ++i) v += c[i] * r[i];
  return v;
}

/* ===================================================================== */
/* loss objects — src/cd_differentiable_function.jl                       */
/* ===================================================================== */
struct cdgpu_handle_s {
  int kind;
  int64_t n, p, ld;
  const double *X; /* n x p (naive) or A p x p (quad); aliased like the Julia structs */
  const double *y; /* y (naive) or b (quad) */
  const double *w; /* WLS only */
  double *state;   /* r (n) or Ax (p) */
  double *ownA, *ownb; /* gram_create owns its A and b */
  double gram_ms;
  sparse_iterate x; /* scratch iterate reused across calls */
  int64_t *order;   /* RandomIterator.order */
  uint64_t *keys;
  uint64_t pass_counter;
};
typedef struct cdgpu_handle_s H;

/* initialize!: :59-72 (LS), :135-148 (WLS), :218-231 (sqrt), :311-320 (quad) */
static void initialize(H *f, const sparse_iterate *x) {
  if (f->kind == CDGPU_LOSS_QUAD) {
    for (int64_t i = 1; i <= f->p; ++i) f->state[i - 1] = A_mul_B_row(f->X, f->ld, x, i);
  } else {
    for (int64_t i = 1; i <= f->n; ++i) f->state[i - 1] = f->y[i - 1] - A_mul_B_row(f->X, f->ld, x, i);
  }
}

/* gradient(f, x, j): :75-76, :150-158, :234-235, :321-322 */
static double gradient(const H *f, int64_t j) {
  switch (f->kind) {
  case CDGPU_LOSS_LS:
    return -At_mul_B_row(f->X, f->ld, f->n, f->state, j) / (double)f->n;
  case CDGPU_LOSS_WLS: {
    double out = 0.0;
    const double *c = f->X + (j - 1) * f->ld;
    for (int64_t i = 0; i < f->n; ++i) out += f->w[i] * c[i] * f->state[i];
    return -out / (double)f->n;
  }
  case CDGPU_LOSS_SQRT: {
    double nr = 0.0;
    for (int64_t i = 0; i < f->n; ++i) nr += f->state[i] * f->state[i];
    return -At_mul_B_row(f->X, f->ld, f->n, f->state, j) / sqrt(nr);
  }
  default:
    return f->state[j - 1] + f->y[j - 1];
  }
}

/* descendCoordinate!(f::CDLeastSquaresLoss, ...) :83-111 */
static double descend_ls(H *f, const prox_l1 *g, sparse_iterate *x, int64_t k) {
  const int64_t n = f->n;
  const double *c = f->X + (k - 1) * f->ld;
  double *r = f->state;
  double a = 0.0, b = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    a += c[i] * c[i];
    b += r[i] * c[i];
  }
  double oldVal = si_get(x, k);
  si_set(x, k, oldVal + b / a);
  double newVal = cdprox(g, x, k, (double)n / a);
  double h = newVal - oldVal;
  for (int64_t i = 0; i < n; ++i) r[i] -= c[i] * h;
  return h;
}
/* descendCoordinate!(f::CDWeightedLSLoss, ...) :165-194 */
static double descend_wls(H *f, const prox_l1 *g, sparse_iterate *x, int64_t k) {
  const int64_t n = f->n;
  const double *c = f->X + (k - 1) * f->ld;
  const double *w = f->w;
  double *r = f->state;
  double a = 0.0, b = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    a += c[i] * c[i] * w[i];
    b += r[i] * c[i] * w[i];
  }
  double oldVal = si_get(x, k);
  si_set(x, k, oldVal + b / a);
  double newVal = cdprox(g, x, k, (double)n / a);
  double h = newVal - oldVal;
  for (int64_t i = 0; i < n; ++i) r[i] -= c[i] * h;
  return h;
}
/* descendCoordinate!(f::CDSqrtLassoLoss, ...) :242-291 */
static double descend_sqrt(H *f, const prox_l1 *g, sparse_iterate *x, int64_t k) {
  const int64_t n = f->n;
  const double *c = f->X + (k - 1) * f->ld;
  double *r = f->state;
  double xk = si_get(x, k);
  for (int64_t i = 0; i < n; ++i) r[i] += c[i] * xk;
  double s = 0.0, xsqr = 0.0, rsqr = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    xsqr += c[i] * c[i];
    s += r[i] * c[i];
    rsqr += r[i] * r[i];
  }
  double lam = g->l0;
  if (g->l) lam *= g->l[k - 1];
  double oldVal = xk, newVal;
  if (fabs(s) <= lam * sqrt(rsqr))
    newVal = 0.0;
  else if (s > lam * sqrt(rsqr))
    newVal = (s - lam / sqrt(1 - lam * lam / xsqr) * sqrt(rsqr - s * s / xsqr)) / xsqr;
  else
    newVal = (s + lam / sqrt(1 - lam * lam / xsqr) * sqrt(rsqr - s * s / xsqr)) / xsqr;
  si_set(x, k, newVal);
  /* note x[k] = v on an absent key with v == 0 stores nothing, as in Julia */
  newVal = si_get(x, k);
  for (int64_t i = 0; i < n; ++i) r[i] -= c[i] * newVal;
  return newVal - oldVal;
}
/* descendCoordinate!(f::CDQuadraticLoss, ...) :324-348 */
static double descend_quad(H *f, const prox_l1 *g, sparse_iterate *x, int64_t k) {
  const int64_t p = f->p;
  const double *A = f->X;
  double a = A[(k - 1) + (k - 1) * f->ld];
  double b = gradient(f, k);
  double oldVal = si_get(x, k);
  a = 1.0 / a;
  si_set(x, k, oldVal - b * a);
  double newVal = cdprox(g, x, k, a);
  double h = newVal - oldVal;
  double *Ax = f->state;
  const double *c = A + (k - 1) * f->ld;
  for (int64_t i = 0; i < p; ++i) Ax[i] += c[i] * h;
  return h;
}
static inline double descend(H *f, const prox_l1 *g, sparse_iterate *x, int64_t k) {
  switch (f->kind) {
  case CDGPU_LOSS_LS:
    return descend_ls(f, g, x, k);
  case CDGPU_LOSS_WLS:
    return descend_wls(f, g, x, k);
  case CDGPU_LOSS_SQRT:
    return descend_sqrt(f, g, x, k);
  default:
    return descend_quad(f, g, x, k);
  }
}

/* ===================================================================== */
/* iterators — src/atom_iterator.jl                                       */
/* ===================================================================== */
/* randomize == 1: permutation = argsort of a counter-based hash (same definition
 * as the device); randomize == 2: the reference's literal Fisher-Yates (:53-64)
 * driven by xoshiro256** in place of Julia's global RNG. */
static inline uint64_t splitmix(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
/* Mode-1 visit order: a keyed bijection of [0, N) — 4-round Feistel network on
 * 2*hb bits with cycle walking — evaluated per position, so neither side needs a
 * sort or a sequential shuffle.  Same definition on the device (csrc/common.cuh). */
static inline uint32_t mix32(uint32_t h) {
  h ^= h >> 16;
  h *= 0x85ebca6bu;
  h ^= h >> 13;
  h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return h;
}
static inline uint64_t order_perm(uint64_t N, uint64_t seed, uint64_t pass, uint64_t x) {
  uint64_t k0 = splitmix(seed ^ (pass * 0xD1B54A32D192ED03ull)), k1 = splitmix(k0);
  uint32_t key[4] = {(uint32_t)k0, (uint32_t)(k0 >> 32), (uint32_t)k1, (uint32_t)(k1 >> 32)};
  int bits = 0;
  while (((N - 1) >> bits) != 0) bits++;
  int hb = (bits + 1) / 2;
  if (hb < 1) hb = 1;
  uint32_t mask = (uint32_t)((1ull << hb) - 1);
  do {
    uint32_t L = (uint32_t)(x >> hb), R = (uint32_t)x & mask;
    for (int r = 0; r < 4; ++r) {
      uint32_t t = L ^ (mix32(R + key[r]) & mask);
      L = R;
      R = t;
    }
    x = ((uint64_t)L << hb) | R;
  } while (x >= N);
  return x;
}
static uint64_t xo_s[4];
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t xo_next(void) {
  uint64_t r = rotl(xo_s[1] * 5, 7) * 9, t = xo_s[1] << 17;
  xo_s[2] ^= xo_s[0];
  xo_s[3] ^= xo_s[1];
  xo_s[1] ^= xo_s[2];
  xo_s[0] ^= xo_s[3];
  xo_s[2] ^= t;
  xo_s[3] = rotl(xo_s[3], 45);
  return r;
}
static void xo_seed(uint64_t seed) {
  for (int i = 0; i < 4; ++i) xo_s[i] = seed = splitmix(seed);
}
static int64_t rand_range(int64_t lo, int64_t hi) { /* rand(lo:hi) */
  uint64_t span = (uint64_t)(hi - lo) + 1, lim = UINT64_MAX - UINT64_MAX % span, v;
  do v = xo_next(); while (v >= lim);
  return lo + (int64_t)(v % span);
}
/* reset!(it, fullPass): :34-37 (ordered), :53-64 (random) */
static void iterator_reset(H *f, const sparse_iterate *x, int fullPass, int randomize, uint64_t seed) {
  int64_t len = fullPass ? f->p : x->nnz;
  if (randomize == 1) {
    for (int64_t i = 0; i < len; ++i)
      f->order[i] = (int64_t)order_perm((uint64_t)len, seed, f->pass_counter, (uint64_t)i) + 1;
  } else if (randomize == 2) {
    for (int64_t i = 1; i <= len; ++i) f->order[i - 1] = i;
    for (int64_t i = 1; i <= len - 1; ++i) {
      int64_t j = rand_range(i, len);
      int64_t t = f->order[i - 1];
      f->order[i - 1] = f->order[j - 1];
      f->order[j - 1] = t;
    }
  }
  f->pass_counter += 1;
}

/* ===================================================================== */
/* driver — src/coordinate_descent.jl                                     */
/* ===================================================================== */
/* _cdPass!: :94-110.  The sparse pass reads nzval2ind live (:25 / :74 of
 * atom_iterator.jl) with the length fixed at the start of the pass? No: the
 * iterator re-evaluates nnz(x) at every step (:21), but no entry can be appended
 * during a sparse pass and zeros stay stored until dropzeros!, so it is constant. */
static double cd_pass(H *f, const prox_l1 *g, sparse_iterate *x, int fullPass, int randomize, cdgpu_stats *st) {
  double maxH = 0.0;
  int64_t i = 1;
  for (;;) {
    int done = fullPass ? i > f->p : i > x->nnz;
    if (done) break;
    int64_t pos = randomize ? f->order[i - 1] : i;
    int64_t k = fullPass ? pos : x->nzval2ind[pos - 1];
    double h = descend(f, g, x, k);
    if (fabs(h) > maxH) maxH = fabs(h);
    st->visits += 1;
    if (h != 0.0) st->accepted += 1;
    i += 1;
  }
  si_dropzeros(x);
  return maxH;
}
/* _coordinateDescent!: :65-92 */
static void cd_loop(H *f, const prox_l1 *g, sparse_iterate *x, const cdgpu_options *o, cdgpu_stats *st) {
  int prev_converged = 0, converged = 1;
  st->converged = 0;
  for (int64_t iter = 1; iter <= o->maxIter; ++iter) {
    iterator_reset(f, x, converged, o->randomize, o->seed);
    st->passes += 1;
    if (converged) st->full_passes += 1;
    double maxH = cd_pass(f, g, x, converged, o->randomize, st);
    st->maxH = maxH;
    prev_converged = converged;
    converged = maxH < o->optTol;
    if (prev_converged && converged) {
      st->converged = 1;
      break;
    }
  }
}
/* _findLambdaMax: :118-131 (scalar), :136-149 (weighted) */
static double find_lambda_max(const H *f, const prox_l1 *g) {
  double lmax = 0.0;
  for (int64_t k = 1; k <= f->p; ++k) {
    double t = fabs(gradient(f, k));
    if (g->l) t = t / g->l[k - 1];
    if (t > lmax) lmax = t;
  }
  return lmax;
}
/* coordinateDescent!(x, f, g::ProxL1, options): :7-39 */
static void coordinate_descent(H *f, const prox_l1 *g, sparse_iterate *x, const cdgpu_options *o, cdgpu_stats *st) {
  if (o->warmStart) {
    initialize(f, x);
    cd_loop(f, g, x, o, st);
  } else {
    si_fill_zero(x);
    initialize(f, x);
    double lmax = find_lambda_max(f, g);
    /* for l in l1:(l2-l1)/numSteps:l2 — numSteps+1 points, the last one l2 */
    double l1 = log(lmax), l2 = log(g->l0);
    double step = (l2 - l1) / (double)o->numSteps;
    for (int64_t i = 0; i <= o->numSteps; ++i) {
      double l = (i == o->numSteps) ? l2 : l1 + (double)i * step;
      prox_l1 g1 = {exp(l), g->l};
      cd_loop(f, &g1, x, o, st);
    }
  }
}

/* ===================================================================== */
/* helpers — src/utils.jl                                                 */
/* ===================================================================== */
/* _stdX!: :127-138 and :140-151 */
static void stdx(const double *X, int64_t n, int64_t p, int64_t ld, const double *w, double *out) {
  for (int64_t j = 0; j < p; ++j) {
    const double *c = X + j * ld;
    double v = 0.0;
    if (w)
      for (int64_t i = 0; i < n; ++i) v += w[i] * (c[i] * c[i]);
    else
      for (int64_t i = 0; i < n; ++i) v += c[i] * c[i];
    out[j] = sqrt(v / (double)n);
  }
}
/* Statistics.std (corrected, two-pass as Julia's varm) */
static double stdev(const double *r, int64_t n) {
  double m = 0.0;
  for (int64_t i = 0; i < n; ++i) m += r[i];
  m /= (double)n;
  double v = 0.0;
  for (int64_t i = 0; i < n; ++i) v += (r[i] - m) * (r[i] - m);
  return sqrt(v / (double)(n - 1));
}
/* least squares Xs \ y by Householder QR (what LAPACK's geqrf path of `\` does
 * for a tall full-rank matrix).  Xs: n x s column-major copy, destroyed. */
static void qr_solve(double *Q, int64_t n, int64_t s, double *rhs, double *beta) {
  for (int64_t j = 0; j < s; ++j) {
    double *c = Q + j * n;
    double nrm = 0.0;
    for (int64_t i = j; i < n; ++i) nrm += c[i] * c[i];
    nrm = sqrt(nrm);
    if (nrm == 0.0) continue;
    double alpha = c[j] > 0 ? -nrm : nrm;
    double v0 = c[j] - alpha;
    double vnorm2 = v0 * v0;
    for (int64_t i = j + 1; i < n; ++i) vnorm2 += c[i] * c[i];
    c[j] = v0;
    /* apply H = I - 2 v v'/v'v to the remaining columns and to rhs */
    for (int64_t jj = j + 1; jj <= s; ++jj) {
      double *t = jj < s ? Q + jj * n : rhs;
      double d = 0.0;
      for (int64_t i = j; i < n; ++i) d += c[i] * t[i];
      d = 2.0 * d / vnorm2;
      for (int64_t i = j; i < n; ++i) t[i] -= d * c[i];
    }
    /* store R's diagonal in place of v0 afterwards: keep alpha aside */
    c[j] = alpha; /* rows > j of column j still hold v (unused later) */
  }
  for (int64_t j = s - 1; j >= 0; --j) {
    double v = rhs[j];
    for (int64_t jj = j + 1; jj < s; ++jj) v -= Q[j + jj * n] * beta[jj];
    beta[j] = v / Q[j + j * n];
  }
}
/* _findLargestCorrelations :96-107 + _findInitResiduals! :66-77 + _findInitSigma! :60-64 */
static int cmp_desc(const void *a, const void *b) {
  double x = *(const double *)a, y = *(const double *)b;
  return x > y ? -1 : (x < y);
}
static int find_init_residuals(const double *X, int64_t n, int64_t p, int64_t ld, const double *y, int64_t s,
                               double *storage) {
  if (s < 1) s = 1;
  if (s > p) s = p;
  double *c = (double *)malloc((size_t)p * sizeof(double));
  double *sorted = (double *)malloc((size_t)p * sizeof(double));
  if (!c || !sorted) return -1;
  for (int64_t j = 1; j <= p; ++j) c[j - 1] = fabs(At_mul_B_row(X, ld, n, y, j));
  memcpy(sorted, c, (size_t)p * sizeof(double));
  qsort(sorted, (size_t)p, sizeof(double), cmp_desc);
  double thr = sorted[s - 1]; /* nlargest(s, storage)[end] */
  int64_t cnt = 0;
  for (int64_t j = 0; j < p; ++j) cnt += c[j] >= thr;
  double *Xs = (double *)malloc((size_t)(n * cnt) * sizeof(double));
  double *Xc = (double *)malloc((size_t)(n * cnt) * sizeof(double));
  double *rhs = (double *)malloc((size_t)n * sizeof(double));
  double *beta = (double *)calloc((size_t)cnt, sizeof(double));
  if (!Xs || !Xc || !rhs || !beta) return -1;
  int64_t q = 0;
  for (int64_t j = 0; j < p; ++j)
    if (c[j] >= thr) {
      memcpy(Xs + q * n, X + j * ld, (size_t)n * sizeof(double));
      q++;
    }
  memcpy(Xc, Xs, (size_t)(n * cnt) * sizeof(double));
  memcpy(rhs, y, (size_t)n * sizeof(double));
  qr_solve(Xc, n, cnt, rhs, beta);
  for (int64_t i = 0; i < n; ++i) {
    double v = 0.0;
    for (int64_t j = 0; j < cnt; ++j) v += Xs[i + j * n] * beta[j];
    storage[i] = y[i] - v;
  }
  free(c);
  free(sorted);
  free(Xs);
  free(Xc);
  free(rhs);
  free(beta);
  return 0;
}

/* ===================================================================== */
/* exported ABI (cdref_*)                                                 */
/* ===================================================================== */
API int cdref_version(void) { return CDGPU_VERSION; }
API const char *cdref_last_error(void) { return g_err; }
API int cdref_device_count(int *count) {
  *count = 0;
  return CDGPU_OK;
}
API void cdref_default_options(cdgpu_options *o) { /* utils.jl:14-20 */
  o->maxIter = 2000;
  o->optTol = 1e-7;
  o->randomize = 1;
  o->warmStart = 1;
  o->numSteps = 50;
  o->seed = 0;
}
API void cdref_default_iter_options(cdgpu_iter_options *o) { /* utils.jl:32-39 */
  o->maxIter = 20;
  o->optTol = 1e-2;
  o->initProcedure = CDGPU_INIT_SCREENING;
  o->_pad = 0;
  o->sinit = 5;
  o->sigma_init = 1.0;
  cdref_default_options(&o->optionsCD);
}

static H *new_handle(int kind, int64_t n, int64_t p, int64_t ld) {
  H *f = (H *)calloc(1, sizeof(H));
  if (!f) return NULL;
  f->kind = kind;
  f->n = n;
  f->p = p;
  f->ld = ld;
  int64_t ns = kind == CDGPU_LOSS_QUAD ? p : n;
  f->state = (double *)calloc((size_t)ns, sizeof(double));
  f->order = (int64_t *)calloc((size_t)p, sizeof(int64_t));
  f->keys = (uint64_t *)calloc((size_t)p, sizeof(uint64_t));
  if (!f->state || !f->order || !f->keys || si_alloc(&f->x, p)) return NULL;
  return f;
}

API int cdref_naive_create(cdgpu_handle *h, int loss_kind, const double *X, int64_t n, int64_t p, int64_t ldx,
                           const double *y, const double *w, int device) {
  (void)device;
  if (!h || !X || !y) return fail(CDGPU_EARG, "null pointer");
  if (loss_kind != CDGPU_LOSS_LS && loss_kind != CDGPU_LOSS_WLS && loss_kind != CDGPU_LOSS_SQRT)
    return fail(CDGPU_EARG, "loss_kind must be LS, WLS or SQRT");
  if ((loss_kind == CDGPU_LOSS_WLS) != (w != NULL)) return fail(CDGPU_EARG, "w must be given iff loss is WLS");
  if (n < 1 || p < 1 || ldx < n) return fail(CDGPU_EDIM, "DimensionMismatch");
  H *f = new_handle(loss_kind, n, p, ldx);
  if (!f) return fail(CDGPU_ENOMEM, "out of memory");
  f->X = X;
  f->y = y;
  f->w = w;
  memcpy(f->state, y, (size_t)n * sizeof(double)); /* r = copy(y) :54 */
  *h = f;
  return CDGPU_OK;
}
API int cdref_naive_create_dev(cdgpu_handle *h, int loss_kind, const double *X, int64_t n, int64_t p, int64_t ldx,
                               const double *y, const double *w, int device) {
  return cdref_naive_create(h, loss_kind, X, n, p, ldx, y, w, device);
}
/* CDQuadraticLoss ctor :305-308 */
API int cdref_quad_create(cdgpu_handle *h, const double *A, int64_t p, int64_t lda, const double *b, int device) {
  (void)device;
  if (!h || !A || !b) return fail(CDGPU_EARG, "null pointer");
  if (p < 1 || lda < p) return fail(CDGPU_EDIM, "DimensionMismatch");
  for (int64_t j = 0; j < p; ++j)
    for (int64_t i = j + 1; i < p; ++i)
      if (A[i + j * lda] != A[j + i * lda]) return fail(CDGPU_EARG, "ArgumentError: A is not symmetric");
  H *f = new_handle(CDGPU_LOSS_QUAD, p, p, lda);
  if (!f) return fail(CDGPU_ENOMEM, "out of memory");
  f->X = A;
  f->y = b;
  *h = f;
  return CDGPU_OK;
}
API int cdref_quad_create_dev(cdgpu_handle *h, const double *A, int64_t p, int64_t lda, const double *b, int device) {
  return cdref_quad_create(h, A, p, lda, b, device);
}
/* A = X'X/n, b = -X'y/n as the reference's users write it (test/lasso.jl:48,88);
 * plain triple loop, lower triangle mirrored so that issymmetric holds. */
API int cdref_gram_create(cdgpu_handle *h, const double *X, int64_t n, int64_t p, int64_t ldx, const double *y,
                          int device) {
  (void)device;
  if (!h || !X || !y) return fail(CDGPU_EARG, "null pointer");
  if (n < 1 || p < 1 || ldx < n) return fail(CDGPU_EDIM, "DimensionMismatch");
  double t0 = now_ms();
  double *A = (double *)malloc((size_t)(p * p) * sizeof(double));
  double *b = (double *)malloc((size_t)p * sizeof(double));
  if (!A || !b) return fail(CDGPU_ENOMEM, "out of memory");
  for (int64_t j = 0; j < p; ++j) {
    const double *cj = X + j * ldx;
    for (int64_t i = j; i < p; ++i) {
      const double *ci = X + i * ldx;
      double v = 0.0;
      for (int64_t k = 0; k < n; ++k) v += ci[k] * cj[k];
      v = v / (double)n;
      A[i + j * p] = v;
      A[j + i * p] = v;
    }
    double v = 0.0;
    for (int64_t k = 0; k < n; ++k) v += cj[k] * y[k];
    b[j] = -v / (double)n;
  }
  int rc = cdref_quad_create(h, A, p, p, b, 0);
  if (rc) {
    free(A);
    free(b);
    return rc;
  }
  (*h)->ownA = A;
  (*h)->ownb = b;
  (*h)->gram_ms = now_ms() - t0;
  return CDGPU_OK;
}
API int cdref_gram_create_dev(cdgpu_handle *h, const double *X, int64_t n, int64_t p, int64_t ldx, const double *y,
                              int device) {
  return cdref_gram_create(h, X, n, p, ldx, y, device);
}
API int cdref_destroy(cdgpu_handle f) {
  if (!f) return CDGPU_OK;
  free(f->state);
  free(f->order);
  free(f->keys);
  free(f->ownA);
  free(f->ownb);
  si_free(&f->x);
  free(f);
  return CDGPU_OK;
}
API int cdref_dims(cdgpu_handle f, int64_t *n, int64_t *p, int *loss_kind) {
  if (!f) return fail(CDGPU_EARG, "null handle");
  if (n) *n = f->n;
  if (p) *p = f->p;
  if (loss_kind) *loss_kind = f->kind;
  return CDGPU_OK;
}
API int cdref_gram_ms(cdgpu_handle f, double *ms) {
  if (!f) return fail(CDGPU_EARG, "null handle");
  *ms = f->gram_ms;
  return CDGPU_OK;
}
API int cdref_quad_get(cdgpu_handle f, double *A_out, double *b_out) {
  if (!f || f->kind != CDGPU_LOSS_QUAD) return fail(CDGPU_EARG, "not a QUAD handle");
  if (A_out)
    for (int64_t j = 0; j < f->p; ++j) memcpy(A_out + j * f->p, f->X + j * f->ld, (size_t)f->p * sizeof(double));
  if (b_out) memcpy(b_out, f->y, (size_t)f->p * sizeof(double));
  return CDGPU_OK;
}

static int check_solve_args(H *f, const cdgpu_options *o) {
  if (!f || !o) return fail(CDGPU_EARG, "null pointer");
  if (o->maxIter < 0 || o->numSteps < 1) return fail(CDGPU_EARG, "bad options");
  return CDGPU_OK;
}

API int cdref_solve(cdgpu_handle f, double lambda0, const double *omega, const cdgpu_options *opt, double *nzval,
                    int64_t *nzval2ind, int64_t *nnz, cdgpu_stats *stats) {
  int rc = check_solve_args(f, opt);
  if (rc) return rc;
  if (!nzval || !nzval2ind || !nnz) return fail(CDGPU_EARG, "null iterate");
  cdgpu_stats st;
  memset(&st, 0, sizeof st);
  double t0 = now_ms();
  if (si_load(&f->x, nzval, nzval2ind, *nnz)) return fail(CDGPU_EDIM, "DimensionMismatch: bad iterate");
  f->pass_counter = 0;
  if (opt->randomize == 2) xo_seed(opt->seed);
  prox_l1 g = {lambda0, omega};
  coordinate_descent(f, &g, &f->x, opt, &st);
  si_store(&f->x, nzval, nzval2ind, nnz);
  st.device_ms = now_ms() - t0;
  if (stats) *stats = st;
  return CDGPU_OK;
}

/* LassoPath ctor loop: lasso.jl:250-257 (on whatever loss the handle holds) */
API int cdref_path(cdgpu_handle f, const double *lambda, int64_t m, const double *omega, const cdgpu_options *opt,
                   int64_t max_hat_s, int64_t capacity, int64_t *colptr, int64_t *rowval, double *nzval,
                   int64_t *m_done, cdgpu_stats *stats) {
  int rc = check_solve_args(f, opt);
  if (rc) return rc;
  if (!lambda || !colptr || !rowval || !nzval || !m_done || m < 0) return fail(CDGPU_EARG, "null pointer");
  si_fill_zero(&f->x);
  f->pass_counter = 0;
  if (opt->randomize == 2) xo_seed(opt->seed);
  colptr[0] = 0;
  *m_done = 0;
  for (int64_t i = 0; i < m; ++i) {
    cdgpu_stats st;
    memset(&st, 0, sizeof st);
    double t0 = now_ms();
    prox_l1 g = {lambda[i], omega};
    coordinate_descent(f, &g, &f->x, opt, &st);
    st.device_ms = now_ms() - t0;
    if (stats) stats[i] = st;
    if (colptr[i] + f->x.nnz > capacity) return fail(CDGPU_ECAP, "path output capacity too small");
    for (int64_t s = 0; s < f->x.nnz; ++s) {
      rowval[colptr[i] + s] = f->x.nzval2ind[s];
      nzval[colptr[i] + s] = f->x.nzval[s];
    }
    colptr[i + 1] = colptr[i] + f->x.nnz;
    *m_done = i + 1;
    if (max_hat_s >= 0 && f->x.nnz > max_hat_s) break;
  }
  return CDGPU_OK;
}

/* scaledLasso!: lasso.jl:107-144 */
API int cdref_scaled_solve(cdgpu_handle f, double lambda, const double *omega, const cdgpu_iter_options *opt,
                           double *nzval, int64_t *nzval2ind, int64_t *nnz, double *sigma_out, cdgpu_stats *stats) {
  if (!f || !opt || !omega || !nzval || !nzval2ind || !nnz) return fail(CDGPU_EARG, "null pointer");
  if (f->kind != CDGPU_LOSS_LS) return fail(CDGPU_EARG, "scaled lasso needs a CDLeastSquaresLoss handle");
  int rc = check_solve_args(f, &opt->optionsCD);
  if (rc) return rc;
  cdgpu_stats st;
  memset(&st, 0, sizeof st);
  double t0 = now_ms();
  if (si_load(&f->x, nzval, nzval2ind, *nnz)) return fail(CDGPU_EDIM, "DimensionMismatch: bad iterate");
  f->pass_counter = 0;
  if (opt->optionsCD.randomize == 2) xo_seed(opt->optionsCD.seed);
  const int64_t n = f->n;
  double sigma;
  if (opt->initProcedure == CDGPU_INIT_SCREENING) {
    if (find_init_residuals(f->X, n, f->p, f->ld, f->y, opt->sinit, f->state)) return fail(CDGPU_ENOMEM, "oom");
    sigma = stdev(f->state, n);
  } else if (opt->initProcedure == CDGPU_INIT_STD) {
    sigma = opt->sigma_init;
  } else if (opt->initProcedure == CDGPU_INIT_WARMSTART) {
    initialize(f, &f->x);
    sigma = stdev(f->state, n);
  } else {
    return fail(CDGPU_EARG, "ArgumentError: Incorrect initialization Symbol");
  }
  prox_l1 g = {lambda * sigma, omega};
  for (int64_t iter = 1; iter <= opt->maxIter; ++iter) {
    st.outer_iters = (int32_t)iter;
    coordinate_descent(f, &g, &f->x, &opt->optionsCD, &st);
    double ss = 0.0;
    for (int64_t i = 0; i < n; ++i) ss += f->state[i] * f->state[i];
    double snew = sqrt(ss / (double)n);
    if (fabs(snew - sigma) / sigma < opt->optTol) break;
    sigma = snew;
    g.l0 = lambda * sigma;
  }
  st.sigma = sigma;
  si_store(&f->x, nzval, nzval2ind, nnz);
  if (sigma_out) *sigma_out = stdev(f->state, n);
  st.device_ms = now_ms() - t0;
  if (stats) *stats = st;
  return CDGPU_OK;
}

API int cdref_state(cdgpu_handle f, double *out) {
  if (!f || !out) return fail(CDGPU_EARG, "null pointer");
  memcpy(out, f->state, (size_t)(f->kind == CDGPU_LOSS_QUAD ? f->p : f->n) * sizeof(double));
  return CDGPU_OK;
}
API int cdref_stdx(cdgpu_handle f, const double *w, double *out) {
  if (!f || !out) return fail(CDGPU_EARG, "null pointer");
  if (f->kind == CDGPU_LOSS_QUAD) { /* A = X'X/n  =>  _stdX!(X)_j = sqrt(A_jj) */
    if (w) return fail(CDGPU_EARG, "weighted stdx needs a naive handle");
    for (int64_t j = 0; j < f->p; ++j) out[j] = sqrt(f->X[j + j * f->ld]);
    return CDGPU_OK;
  }
  stdx(f->X, f->n, f->p, f->ld, w, out);
  return CDGPU_OK;
}
API int cdref_lambda_max(cdgpu_handle f, const double *omega, double *out) {
  if (!f || !out) return fail(CDGPU_EARG, "null pointer");
  si_fill_zero(&f->x);
  initialize(f, &f->x);
  prox_l1 g = {0.0, omega};
  *out = find_lambda_max(f, &g);
  return CDGPU_OK;
}

/* ---------------------------------------------------------------------- */
/* varying-coefficient lasso — src/varying_coefficient_lasso.jl           */
/* ---------------------------------------------------------------------- */
/* evaluate: :17-21 */
API double cdref_kernel_evaluate(int kind, double h, double x, double y) {
  if (kind == CDGPU_KERNEL_GAUSSIAN) return exp(-((x - y) * (x - y)) / h) / h;
  double u = (x - y) / h;
  return fabs(u) >= 1.0 ? 0.0 : 0.75 * (1.0 - u * u) / h;
}
/* _expand_X!: :550-569.  tX is n x p*(degree+1), ld n */
API void cdref_expand_X(double *tX, const double *X, int64_t n, int64_t p, int64_t ldx, const double *z, double z0,
                        int degree) {
  for (int64_t j = 0; j < p; ++j)
    for (int64_t i = 0; i < n; ++i) {
      double v = X[i + j * ldx];
      double df = z[i] - z0;
      int64_t col = j * (degree + 1);
      tX[i + col * n] = v;
      for (int l = 1; l <= degree; ++l) {
        v *= df;
        tX[i + (col + l) * n] = v;
      }
    }
}
/* The iterate is carried from one grid point to the next exactly as in the reference (:56,:68). */
/* A x = rhs for a dense k x k system by LU with partial pivoting (what Julia's `\` does for a general square
 * matrix); A (column-major, ld k) and rhs are overwritten.  Returns 1 on a zero pivot. */
static int lu_solve(double *A, double *rhs, int64_t k) {
  for (int64_t c = 0; c < k; ++c) {
    int64_t piv = c;
    for (int64_t i = c + 1; i < k; ++i)
      if (fabs(A[i + c * k]) > fabs(A[piv + c * k])) piv = i;
    if (A[piv + c * k] == 0.0) return 1;
    if (piv != c) {
      for (int64_t j = 0; j < k; ++j) {
        double t = A[c + j * k];
        A[c + j * k] = A[piv + j * k];
        A[piv + j * k] = t;
      }
      double t = rhs[c];
      rhs[c] = rhs[piv];
      rhs[piv] = t;
    }
    for (int64_t i = c + 1; i < k; ++i) {
      double l = A[i + c * k] / A[c + c * k];
      A[i + c * k] = l;
      for (int64_t j = c + 1; j < k; ++j) A[i + j * k] -= l * A[c + j * k];
      rhs[i] -= l * rhs[c];
    }
  }
  for (int64_t c = k - 1; c >= 0; --c) {
    double v = rhs[c];
    for (int64_t j = c + 1; j < k; ++j) v -= A[c + j * k] * rhs[j];
    rhs[c] = v / A[c + c * k];
  }
  return 0;
}

/* locpolyl1 (:30-79) with the optional refit (:71-76): outR[S, g] = (Xs' W Xs) \ (Xs' W y) on the expanded
 * coordinates S of every group with a non-zero coefficient (get_nonzero_coordinates!, :488-512).  outR may be
 * NULL (refit = false). */
API int cdref_vc_solve_refit(const double *X, int64_t n, int64_t p, int64_t ldx, const double *z, const double *y,
                             const double *zgrid, int64_t m, int64_t m_begin, int64_t m_end, int degree, int kernel_kind,
                             double bandwidth, double lambda0, const cdgpu_options *opt, int device, double *out,
                             double *outR, cdgpu_stats *stats) {
  (void)device;
  if (!X || !z || !y || !zgrid || !opt || !out) return fail(CDGPU_EARG, "null pointer");
  if (n < 1 || p < 1 || ldx < n || degree < 0 || m < 0 || m_begin < 0 || m_end > m || m_begin > m_end)
    return fail(CDGPU_EDIM, "DimensionMismatch");
  if (kernel_kind != CDGPU_KERNEL_GAUSSIAN && kernel_kind != CDGPU_KERNEL_EPANECHNIKOV)
    return fail(CDGPU_EARG, "unknown kernel");
  const int64_t ep = p * (degree + 1);
  double *w = (double *)malloc((size_t)n * sizeof(double));
  double *eX = (double *)calloc((size_t)(n * ep), sizeof(double));
  double *sd = (double *)malloc((size_t)ep * sizeof(double));
  if (!w || !eX || !sd) return fail(CDGPU_ENOMEM, "out of memory");
  cdgpu_handle f;
  /* CDWeightedLSLoss(y, expandX, w): buffers aliased and refilled per grid point (:54) */
  for (int64_t i = 0; i < n; ++i) w[i] = 1.0;
  int rc = cdref_naive_create(&f, CDGPU_LOSS_WLS, eX, n, ep, n, y, w, 0);
  if (rc) return rc;
  cdgpu_options o = *opt;
  o.warmStart = 1; /* :42 */
  if (o.randomize == 2) xo_seed(o.seed);
  si_fill_zero(&f->x);
  for (int64_t g = m_begin; g < m_end; ++g) {
    double z0 = zgrid[g];
    cdgpu_stats st;
    memset(&st, 0, sizeof st);
    double t0 = now_ms();
    for (int64_t i = 0; i < n; ++i) w[i] = cdref_kernel_evaluate(kernel_kind, bandwidth, z[i], z0);
    cdref_expand_X(eX, X, n, p, ldx, z, z0, degree);
    stdx(eX, n, ep, n, w, sd);
    prox_l1 pen = {lambda0, sd};
    f->pass_counter = 0;
    coordinate_descent(f, &pen, &f->x, &o, &st);
    double *col = out + g * ep;
    for (int64_t k = 0; k < ep; ++k) col[k] = 0.0;
    for (int64_t s = 0; s < f->x.nnz; ++s) col[f->x.nzval2ind[s] - 1] = f->x.nzval[s];
    st.device_ms = now_ms() - t0;
    if (stats) stats[g] = st;
    if (outR) {
      double *colR = outR + g * ep;
      int64_t *S = (int64_t *)malloc((size_t)ep * sizeof(int64_t)), ns = 0, dgp = degree + 1;
      for (int64_t k = 0; k < ep; ++k) colR[k] = 0.0;
      for (int64_t j = 0; j < p; ++j) {
        int nz = 0;
        for (int64_t k = j * dgp; k < (j + 1) * dgp; ++k) nz |= col[k] != 0.0;
        if (nz)
          for (int64_t k = j * dgp; k < (j + 1) * dgp; ++k) S[ns++] = k;
      }
      if (ns > 0) {
        double *M = (double *)malloc((size_t)(ns * ns) * sizeof(double)), *rhs = (double *)malloc((size_t)ns * sizeof(double));
        for (int64_t a = 0; a < ns; ++a) {
          const double *ca = eX + S[a] * n;
          double r = 0.0;
          for (int64_t i = 0; i < n; ++i) r += ca[i] * w[i] * y[i];
          rhs[a] = r;
          for (int64_t b = 0; b < ns; ++b) {
            const double *cb = eX + S[b] * n;
            double v = 0.0;
            for (int64_t i = 0; i < n; ++i) v += ca[i] * w[i] * cb[i];
            M[a + b * ns] = v;
          }
        }
        if (lu_solve(M, rhs, ns))
          for (int64_t a = 0; a < ns; ++a) rhs[a] = NAN; /* SingularException in the reference */
        for (int64_t a = 0; a < ns; ++a) colR[S[a]] = rhs[a];
        free(M);
        free(rhs);
      }
      free(S);
    }
  }
  cdref_destroy(f);
  free(w);
  free(eX);
  free(sd);
  return CDGPU_OK;
}
API int cdref_vc_solve(const double *X, int64_t n, int64_t p, int64_t ldx, const double *z, const double *y,
                       const double *zgrid, int64_t m, int64_t m_begin, int64_t m_end, int degree, int kernel_kind,
                       double bandwidth, double lambda0, const cdgpu_options *opt, int device, double *out,
                       cdgpu_stats *stats) {
  return cdref_vc_solve_refit(X, n, p, ldx, z, y, zgrid, m, m_begin, m_end, degree, kernel_kind, bandwidth, lambda0, opt,
                              device, out, NULL, stats);
}

/* the chain cut into runs of `chain` grid points (include/cdgpu.h: cdgpu_vc_solve_chain): every run is the
 * reference's loop on its own grid points, the first one from zero */
API int cdref_vc_solve_chain(const double *X, int64_t n, int64_t p, int64_t ldx, const double *z, const double *y,
                             const double *zgrid, int64_t m, int64_t m_begin, int64_t m_end, int degree, int kernel_kind,
                             double bandwidth, double lambda0, const cdgpu_options *opt, int64_t chain, int device,
                             double *out, double *outR, cdgpu_stats *stats) {
  if (chain < 1) return fail(CDGPU_EARG, "chain must be at least 1");
  if (m_begin < 0 || m_end > m || m_begin > m_end) return fail(CDGPU_EDIM, "DimensionMismatch");
  for (int64_t b0 = m_begin; b0 < m_end; b0 += chain) {
    const int64_t b1 = b0 + chain < m_end ? b0 + chain : m_end;
    int rc = cdref_vc_solve_refit(X, n, p, ldx, z, y, zgrid, m, b0, b1, degree, kernel_kind, bandwidth, lambda0, opt, device,
                                  out, outR, stats);
    if (rc) return rc;
  }
  return CDGPU_OK;
}

/* refitLassoPath (lasso.jl:208-225) for one support: X[:, S] \ y by the normal equations and LU (the reference's
 * `\` on a tall matrix is a QR least squares: same solution up to conditioning). */
API int cdref_refit(cdgpu_handle f, const int64_t *support, int64_t ns, double *coef_out) {
  if (!f || (ns > 0 && (!support || !coef_out))) return fail(CDGPU_EARG, "null pointer");
  if (ns < 0 || ns > f->p) return fail(CDGPU_EDIM, "DimensionMismatch");
  if (ns == 0) return CDGPU_OK;
  for (int64_t a = 0; a < ns; ++a)
    if (support[a] < 1 || support[a] > f->p) return fail(CDGPU_EDIM, "BoundsError: support index out of range");
  double *M = (double *)malloc((size_t)(ns * ns) * sizeof(double));
  if (!M) return fail(CDGPU_ENOMEM, "out of memory");
  if (f->kind == CDGPU_LOSS_QUAD) {
    for (int64_t a = 0; a < ns; ++a) {
      coef_out[a] = -f->y[support[a] - 1];
      for (int64_t b = 0; b < ns; ++b) M[a + b * ns] = f->X[(support[a] - 1) + (support[b] - 1) * f->ld];
    }
  } else {
    for (int64_t a = 0; a < ns; ++a) {
      const double *ca = f->X + (support[a] - 1) * f->ld;
      double r = 0.0;
      for (int64_t i = 0; i < f->n; ++i) r += ca[i] * (f->kind == CDGPU_LOSS_WLS ? f->w[i] : 1.0) * f->y[i];
      coef_out[a] = r;
      for (int64_t b = 0; b < ns; ++b) {
        const double *cb = f->X + (support[b] - 1) * f->ld;
        double v = 0.0;
        for (int64_t i = 0; i < f->n; ++i) v += ca[i] * (f->kind == CDGPU_LOSS_WLS ? f->w[i] : 1.0) * cb[i];
        M[a + b * ns] = v;
      }
    }
  }
  int singular = lu_solve(M, coef_out, ns);
  free(M);
  if (singular) return fail(CDGPU_EARG, "SingularException: the selected columns are not linearly independent");
  return CDGPU_OK;
}

/* lvocv_locpolyl1 (varying_coefficient_lasso.jl:81-137): for every bandwidth and every observation i the
 * leave-one-out local problem at z0 = z_i (w_i = 0), sigma initialised from the residuals of a weighted LS on the
 * min(10, ep) most correlated columns (utils.jl:79-92, 107-122), <= 10 rounds of CD at lambda0*sigma with
 * sigma = _getSigma(w, r) until it moves by < 1e-2 (:112-124), refit on the selected groups and the squared error
 * of the prediction of y_i (:127-131).  sqerr[q] for problem q = ih * n + i, q in [q_begin, q_end); the caller sums
 * MSE[ih] over i.  opt->warmStart != 0: beta is carried from problem to problem as in the reference (:99,:116);
 * opt->warmStart == 0: every problem starts from beta = 0 (what the batched device path does). */
API int cdref_vc_lvocv(const double *X, int64_t n, int64_t p, int64_t ldx, const double *z, const double *y, int degree,
                       const double *hArr, int64_t numH, int kernel_kind, double lambda0, const cdgpu_options *opt,
                       int64_t q_begin, int64_t q_end, int device, double *sqerr, cdgpu_stats *stats) {
  (void)device;
  if (!X || !z || !y || !hArr || !opt || !sqerr) return fail(CDGPU_EARG, "null pointer");
  const int64_t m = numH * n;
  if (n < 2 || p < 1 || ldx < n || degree < 0 || numH < 0 || q_begin < 0 || q_end > m || q_begin > q_end)
    return fail(CDGPU_EDIM, "DimensionMismatch");
  if (kernel_kind != CDGPU_KERNEL_GAUSSIAN && kernel_kind != CDGPU_KERNEL_EPANECHNIKOV)
    return fail(CDGPU_EARG, "unknown kernel");
  const int64_t ep = p * (degree + 1), dgp = degree + 1;
  double *w = (double *)malloc((size_t)n * sizeof(double));
  double *eX = (double *)calloc((size_t)(n * ep), sizeof(double));
  double *sd = (double *)malloc((size_t)ep * sizeof(double)), *cor = (double *)malloc((size_t)ep * sizeof(double));
  double *srt = (double *)malloc((size_t)ep * sizeof(double)), *M = (double *)malloc((size_t)(ep * ep) * sizeof(double));
  double *rhs = (double *)malloc((size_t)ep * sizeof(double)), *beta = (double *)malloc((size_t)ep * sizeof(double));
  int64_t *S = (int64_t *)malloc((size_t)ep * sizeof(int64_t));
  if (!w || !eX || !sd || !cor || !srt || !M || !rhs || !beta || !S) return fail(CDGPU_ENOMEM, "out of memory");
  for (int64_t i = 0; i < n; ++i) w[i] = 1.0;
  cdgpu_handle f;
  int rc = cdref_naive_create(&f, CDGPU_LOSS_WLS, eX, n, ep, n, y, w, 0);
  if (rc) return rc;
  cdgpu_options o = *opt;
  o.warmStart = 1; /* :92 */
  if (o.randomize == 2) xo_seed(o.seed);
  si_fill_zero(&f->x);
  /* weighted normal equations on the ns columns S of eX: solution into rhs; returns 1 when singular */
#define NORMAL_EQ_SOLVE(ns)                                              \
  do {                                                                   \
    for (int64_t a_ = 0; a_ < (ns); ++a_) {                              \
      const double *ca = eX + S[a_] * n;                                 \
      double r_ = 0.0;                                                   \
      for (int64_t i_ = 0; i_ < n; ++i_) r_ += ca[i_] * w[i_] * y[i_];   \
      rhs[a_] = r_;                                                      \
      for (int64_t b_ = 0; b_ < (ns); ++b_) {                            \
        const double *cb = eX + S[b_] * n;                               \
        double v_ = 0.0;                                                 \
        for (int64_t i_ = 0; i_ < n; ++i_) v_ += ca[i_] * w[i_] * cb[i_]; \
        M[a_ + b_ * (ns)] = v_;                                          \
      }                                                                  \
    }                                                                    \
    singular = lu_solve(M, rhs, (ns));                                   \
  } while (0)
  for (int64_t q = q_begin; q < q_end; ++q) {
    const int64_t ih = q / n, io = q % n;
    const double z0 = z[io], bw = hArr[ih];
    cdgpu_stats st;
    memset(&st, 0, sizeof st);
    double t0 = now_ms();
    int singular = 0;
    if (!opt->warmStart) si_fill_zero(&f->x);
    for (int64_t i = 0; i < n; ++i) w[i] = cdref_kernel_evaluate(kernel_kind, bw, z[i], z0);
    w[io] = 0.0; /* :108 */
    cdref_expand_X(eX, X, n, p, ldx, z, z0, degree);
    stdx(eX, n, ep, n, w, sd);
    /* _findLargestCorrelations(w, X, y, s), utils.jl:107-122: S = storage .>= nlargest(s, storage)[end] */
    const int64_t s_init = ep < 10 ? ep : 10;
    for (int64_t k = 0; k < ep; ++k) {
      const double *ck = eX + k * n;
      double v = 0.0;
      for (int64_t i = 0; i < n; ++i) v += ck[i] * w[i] * y[i];
      cor[k] = srt[k] = fabs(v);
    }
    for (int64_t a_ = 0; a_ < s_init; ++a_) { /* partial selection sort, descending */
      int64_t best = a_;
      for (int64_t b_ = a_ + 1; b_ < ep; ++b_)
        if (srt[b_] > srt[best]) best = b_;
      double t = srt[a_];
      srt[a_] = srt[best];
      srt[best] = t;
    }
    int64_t ns = 0;
    for (int64_t k = 0; k < ep; ++k)
      if (cor[k] >= srt[s_init - 1]) S[ns++] = k;
    NORMAL_EQ_SOLVE(ns);
    double swr = 0.0, sw = 0.0;
    for (int64_t i = 0; i < n; ++i) { /* residuals of that fit, _getSigma (utils.jl:167-175) */
      double fit = 0.0;
      for (int64_t a_ = 0; a_ < ns; ++a_) fit += eX[i + S[a_] * n] * rhs[a_];
      const double r = y[i] - fit;
      swr += r * r * w[i];
      sw += w[i];
    }
    double sigma = sqrt(swr / sw);
    for (int outer = 1; outer <= 10; ++outer) { /* :114-124 */
      prox_l1 pen = {lambda0 * sigma, sd};
      f->pass_counter = 0;
      coordinate_descent(f, &pen, &f->x, &o, &st);
      st.outer_iters = outer;
      swr = 0.0;
      for (int64_t i = 0; i < n; ++i) swr += f->state[i] * f->state[i] * w[i];
      const double snew = sqrt(swr / sw);
      if (fabs(snew - sigma) / sigma < 1e-2) break;
      sigma = snew;
    }
    st.sigma = sigma;
    /* refit on the selected groups and prediction of the left-out response, :127-131 */
    for (int64_t k = 0; k < ep; ++k) beta[k] = 0.0;
    for (int64_t s_ = 0; s_ < f->x.nnz; ++s_) beta[f->x.nzval2ind[s_] - 1] = f->x.nzval[s_];
    ns = 0;
    for (int64_t j = 0; j < p; ++j) {
      int nz = 0;
      for (int64_t k = j * dgp; k < (j + 1) * dgp; ++k) nz |= beta[k] != 0.0;
      if (nz)
        for (int64_t k = j * dgp; k < (j + 1) * dgp; ++k) S[ns++] = k;
    }
    double yh = 0.0;
    if (ns > 0) {
      NORMAL_EQ_SOLVE(ns);
      for (int64_t a_ = 0; a_ < ns; ++a_) yh += eX[io + S[a_] * n] * rhs[a_];
    }
    sqerr[q] = singular ? NAN : (yh - y[io]) * (yh - y[io]);
    st.device_ms = now_ms() - t0;
    if (stats) stats[q] = st;
  }
#undef NORMAL_EQ_SOLVE
  cdref_destroy(f);
  free(w); free(eX); free(sd); free(cor); free(srt); free(M); free(rhs); free(beta); free(S);
  return CDGPU_OK;
}

/* ---------------------------------------------------------------------- */
/* unit-test hooks for the pieces the reference tests directly             */
/* ---------------------------------------------------------------------- */
/* Replays a sequence of x[k] = v stores on an empty SparseIterate(p), optionally
 * runs dropzeros!, and returns the stored triple (test/atom_iterator.jl:11-28). */
API int cdref_sparse_iterate_replay(int64_t p, const int64_t *keys, const double *vals, int64_t nops, int dropzeros,
                                    double *nzval, int64_t *nzval2ind, int64_t *nnz) {
  sparse_iterate x;
  if (si_alloc(&x, p)) return fail(CDGPU_ENOMEM, "oom");
  for (int64_t i = 0; i < nops; ++i) {
    if (keys[i] < 1 || keys[i] > p) {
      si_free(&x);
      return fail(CDGPU_EDIM, "BoundsError");
    }
    si_set(&x, keys[i], vals[i]);
  }
  if (dropzeros) si_dropzeros(&x);
  si_store(&x, nzval, nzval2ind, nnz);
  si_free(&x);
  return CDGPU_OK;
}
/* collect(it) for an Ordered/Random iterator over the given iterate
 * (test/atom_iterator.jl:20-28,56-66): writes the visited coordinates. */
API int cdref_iterator_collect(int64_t p, const int64_t *nzval2ind, int64_t nnz, int fullPass, int randomize,
                               uint64_t seed, uint64_t pass_counter, int64_t *visited, int64_t *order_out) {
  H f;
  memset(&f, 0, sizeof f);
  f.p = p;
  f.order = (int64_t *)calloc((size_t)p, sizeof(int64_t));
  f.keys = (uint64_t *)calloc((size_t)p, sizeof(uint64_t));
  f.pass_counter = pass_counter;
  sparse_iterate x;
  memset(&x, 0, sizeof x);
  x.p = p;
  x.nnz = nnz;
  x.nzval2ind = (int64_t *)nzval2ind;
  if (randomize == 2) xo_seed(seed);
  iterator_reset(&f, &x, fullPass, randomize, seed);
  int64_t len = fullPass ? p : nnz;
  for (int64_t i = 1; i <= len; ++i) {
    int64_t pos = randomize ? f.order[i - 1] : i;
    visited[i - 1] = fullPass ? pos : nzval2ind[pos - 1];
    if (order_out) order_out[i - 1] = pos;
  }
  free(f.order);
  free(f.keys);
  return CDGPU_OK;
}
API double cdref_shrink(double v, double c) { return shrink(v, c); }
API uint64_t cdref_order_perm(uint64_t N, uint64_t seed, uint64_t pass, uint64_t x) { return order_perm(N, seed, pass, x); }
