/*
 * cdgpu.h — C ABI of libcdgpu.so, the B200 (sm_100a) drop-in for the active-set
 * proximal coordinate-descent hot path of CoordinateDescent.jl.
 *
 * The reference has no FFI layer; its "operator API" is the four-function loss
 * protocol initialize!/gradient/numCoordinates/descendCoordinate!
 * (src/cd_differentiable_function.jl:1-35) called once per coordinate from
 * _cdPass! (src/coordinate_descent.jl:94-110).  A per-coordinate FFI call is
 * useless for a GPU, so this ABI sits one level up: one call == one whole
 * coordinateDescent! / LassoPath / scaledLasso! / locpolyl1 invocation.
 *
 * Conventions
 *   - every entry point returns an int status (CDGPU_OK == 0); the message for a
 *     non-zero status is cdgpu_last_error() (thread-local).
 *   - all matrices are column-major Float64 with an explicit leading dimension,
 *     exactly Julia's Matrix{Float64} / StridedMatrix layout.
 *   - the iterate crosses the boundary as ProximalBase's SparseIterate triple
 *     (nzval, nzval2ind, nnz): nzval2ind is 1-BASED Int64, both arrays have
 *     capacity p, entries 1..nnz are meaningful, order == visit order of an
 *     active-set pass (test/atom_iterator.jl:13-28).
 *   - host pointers are owned by the caller; the library copies what it needs to
 *     the device at *_create time and owns all device memory behind the handle.
 *   - calls are blocking; a handle is not thread-safe; no C++ exception crosses.
 *   - there is NO CPU fallback: without a usable CUDA device every compute entry
 *     point fails with CDGPU_ENODEV.
 *
 * The CPU oracle (oracle/cdref.c -> libcdref.so) exports the same functions with
 * the prefix cdref_ instead of cdgpu_ so one harness can drive both.
 */
#ifndef CDGPU_H
#define CDGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CDGPU_VERSION 100 /* 0.1.0 */

/* status codes; mapping to the reference's exceptions in parentheses */
enum {
  CDGPU_OK = 0,
  CDGPU_EDIM = 1,   /* DimensionMismatch: coordinate_descent.jl:13,15; cd_differentiable_function.jl:53,129,212 */
  CDGPU_EARG = 2,   /* ArgumentError: cd_differentiable_function.jl:306; lasso.jl:128 */
  CDGPU_ECUDA = 3,  /* ErrorException(cdgpu_last_error()) */
  CDGPU_ENOMEM = 4, /* ErrorException */
  CDGPU_ENCCL = 5,  /* ErrorException */
  CDGPU_ENODEV = 6, /* ErrorException: no CUDA device / driver (no CPU fallback) */
  CDGPU_ECAP = 7    /* caller-provided output capacity too small */
};

/* loss kinds == the four concrete CoordinateDifferentiableFunction types */
enum {
  CDGPU_LOSS_LS = 0,   /* CDLeastSquaresLoss  cd_differentiable_function.jl:43-111  */
  CDGPU_LOSS_WLS = 1,  /* CDWeightedLSLoss    cd_differentiable_function.jl:118-194 */
  CDGPU_LOSS_SQRT = 2, /* CDSqrtLassoLoss     cd_differentiable_function.jl:202-291 */
  CDGPU_LOSS_QUAD = 3  /* CDQuadraticLoss     cd_differentiable_function.jl:299-348 */
};

/* randomize: 0 = OrderedIterator (atom_iterator.jl:9-37)
 *            1 = RandomIterator  (atom_iterator.jl:41-75): a fresh uniform random
 *                permutation per pass.  Julia's global RNG stream cannot be
 *                reproduced; the permutation is the argsort of a counter-based
 *                hash keyed by (seed, pass counter, position), identical in the
 *                oracle (mode 1) and on the device. */
typedef struct cdgpu_options {
  int64_t maxIter;   /* utils.jl:8   default 2000 */
  double optTol;     /* utils.jl:9   default 1e-7 */
  int32_t randomize; /* utils.jl:10  default 1 (true) */
  int32_t warmStart; /* utils.jl:11  default 1 (true) */
  int64_t numSteps;  /* utils.jl:12  default 50 */
  uint64_t seed;     /* replaces Julia's global RNG; default 0 */
} cdgpu_options;

/* initProcedure of IterLassoOptions (utils.jl:24-39) */
enum { CDGPU_INIT_SCREENING = 0, CDGPU_INIT_STD = 1, CDGPU_INIT_WARMSTART = 2 };

typedef struct cdgpu_iter_options {
  int64_t maxIter;       /* utils.jl:25 default 20   */
  double optTol;         /* utils.jl:26 default 1e-2 */
  int32_t initProcedure; /* utils.jl:27 default :Screening */
  int32_t _pad;
  int64_t sinit;         /* utils.jl:28 default 5  */
  double sigma_init;     /* utils.jl:29 default 1. */
  cdgpu_options optionsCD;
} cdgpu_iter_options;

/* What one solve did.  The reference keeps no such record (it is silent even on
 * hitting maxIter, coordinate_descent.jl:74-91); the counters are what the
 * "coordinate updates / s" metric is computed from. */
typedef struct cdgpu_stats {
  int64_t passes;      /* _cdPass! calls                        */
  int64_t full_passes; /* of which over all p coordinates       */
  int64_t visits;      /* descendCoordinate! calls              */
  int64_t accepted;    /* visits with h != 0                    */
  double maxH;         /* max|h| of the last pass               */
  int32_t converged;   /* 1 iff the loop left through :89       */
  int32_t outer_iters; /* sigma iterations (scaled lasso), else 0 */
  double sigma;        /* scaled lasso: last sigma used; else 0 */
  double device_ms;    /* device time of the solve, CUDA events on the library stream (oracle: host wall ms) */
} cdgpu_stats;

typedef struct cdgpu_handle_s *cdgpu_handle; /* the loss object "f" */

/* smoothing kernels of varying_coefficient_lasso.jl:5-21 */
enum { CDGPU_KERNEL_GAUSSIAN = 0, CDGPU_KERNEL_EPANECHNIKOV = 1 };

/* ---------------------------------------------------------------- misc -- */
int cdgpu_version(void);
const char *cdgpu_last_error(void);
int cdgpu_device_count(int *count);
int cdgpu_launch_count(int64_t *count); /* kernels launched by this library so far (diagnostic) */
void cdgpu_default_options(cdgpu_options *o);           /* CDOptions()          utils.jl:14-20 */
void cdgpu_default_iter_options(cdgpu_iter_options *o); /* IterLassoOptions()   utils.jl:32-39 */

/* ------------------------------------------------------------- handles -- */
/* CDLeastSquaresLoss(y,X) / CDWeightedLSLoss(y,X,w) / CDSqrtLassoLoss(y,X)
 * (cd_differentiable_function.jl:52-55,128-131,211-214).  loss_kind in
 * {LS,WLS,SQRT}; w must be non-NULL iff WLS.  n,p >= 1, ldx >= n else EDIM.
 * X, y, w are HOST pointers and are copied to device `device`. */
int cdgpu_naive_create(cdgpu_handle *h, int loss_kind, const double *X, int64_t n, int64_t p, int64_t ldx,
                       const double *y, const double *w, int device);
/* same, X/y/w are DEVICE pointers on `device`; X is used in place (not copied,
 * never written) and must outlive the handle. */
int cdgpu_naive_create_dev(cdgpu_handle *h, int loss_kind, const double *dX, int64_t n, int64_t p, int64_t ldx,
                           const double *dy, const double *dw, int device);

/* CDQuadraticLoss(A,b) (cd_differentiable_function.jl:305-308): A must be exactly
 * symmetric and length(b)==size(A,2) else EARG. */
int cdgpu_quad_create(cdgpu_handle *h, const double *A, int64_t p, int64_t lda, const double *b, int device);
int cdgpu_quad_create_dev(cdgpu_handle *h, const double *dA, int64_t p, int64_t lda, const double *db, int device);

/* Covariance form straight from the data: A = X'X/n, b = -X'y/n (what the
 * reference's users write by hand, test/lasso.jl:48,88) formed on the device by
 * the FP64 tensor-core SYRK kernel, then wrapped as a QUAD handle. */
int cdgpu_gram_create(cdgpu_handle *h, const double *X, int64_t n, int64_t p, int64_t ldx, const double *y,
                      int device);
int cdgpu_gram_create_dev(cdgpu_handle *h, const double *dX, int64_t n, int64_t p, int64_t ldx, const double *dy,
                          int device);
/* LAZY covariance form: the same CDQuadraticLoss(X'X/n, -X'y/n) object, but only diag(A) and b are formed up front
 * (one pass over X); a column A[:,k] = X'X[:,k]/n is formed the first time coordinate k becomes non-zero, in batches
 * of 128 columns through the same FP64 tensor-core kernel (a skinny GEMM), and cached.  The CD step of
 * cd_differentiable_function.jl:324-348 reads nothing else, so iterates agree with the eager handle to rounding; a path
 * whose supports stay small never pays for the p x p matrix.  The handle keeps X resident on the device (the _dev form
 * uses the caller's buffer in place: it must outlive the handle).  cdgpu_quad_get(A) forms the full matrix on request.
 * Capacity: 4096 cached columns (CDGPU_LAZY_CAP); a solve whose active set outgrows it fails with CDGPU_ECAP — use the
 * eager form for dense solutions. */
int cdgpu_gram_create_lazy(cdgpu_handle *h, const double *X, int64_t n, int64_t p, int64_t ldx, const double *y,
                           int device);
int cdgpu_gram_create_lazy_dev(cdgpu_handle *h, const double *dX, int64_t n, int64_t p, int64_t ldx, const double *dy,
                               int device);
/* what the last solve on a lazy handle formed: columns cached so far, batches formed and kernel pauses during the last
 * solve, device ms spent forming columns during the last solve (all 0 for other handles) */
int cdgpu_lazy_stats(cdgpu_handle h, int64_t *columns, int64_t *batches, int64_t *pauses, double *form_ms);
/* columns formed by the BLOCKING batches of the last solve, i.e. the ones form_ms of cdgpu_lazy_stats times (batches
 * formed in the background, while the sweep kernel runs, are not timed) */
int cdgpu_lazy_form_columns(cdgpu_handle h, int64_t *columns);
/* device ms (CUDA events) spent inside the covariance-form sweep kernel during the last solve / path on a QUAD handle
 * (all its launches; excludes forming columns on a lazy handle) */
int cdgpu_sweep_ms(cdgpu_handle h, double *ms);
/* Row-sharded variant for one-process-per-GPU jobs: every rank passes its own
 * n_local rows; partial X'X and X'y are summed over ranks with one
 * ncclAllReduce on the communicator made by cdgpu_comm_init (NULL comm == single
 * rank), then scaled by 1/n_total.  Every rank ends with the same QUAD handle. */
typedef struct cdgpu_comm_s *cdgpu_comm;
int cdgpu_comm_unique_id(void *id128);                    /* rank 0: 128-byte ncclUniqueId */
int cdgpu_comm_init(cdgpu_comm *c, const void *id128, int rank, int nranks, int device);
int cdgpu_comm_destroy(cdgpu_comm c);
int cdgpu_gram_create_sharded(cdgpu_handle *h, const double *dX_local, int64_t n_local, int64_t n_total, int64_t p,
                              int64_t ldx, const double *dy_local, cdgpu_comm comm, int device);

/* Synthetic design for benchmarks / multi-GPU tests: d_out[i + j*ld] (device memory, rows x cols) = a reproducible
 * pseudo-normal value that depends only on (seed, row0 + i, col0 + j) — every row / column sharding of the same matrix
 * over any number of GPUs sees identical data.  Blocking. */
int cdgpu_synth_normal(double *d_out, int64_t rows, int64_t cols, int64_t ld, int64_t row0, int64_t col0, uint64_t seed,
                       int device);

int cdgpu_destroy(cdgpu_handle h);
int cdgpu_dims(cdgpu_handle h, int64_t *n, int64_t *p, int *loss_kind);
/* device time (ms, CUDA events) of the last Gram formation behind a QUAD handle
 * made by cdgpu_gram_create*; 0 otherwise */
int cdgpu_gram_ms(cdgpu_handle h, double *ms);
/* copy A (p*p, ld p) and/or b back to the host (Julia's f.A / f.b) */
int cdgpu_quad_get(cdgpu_handle h, double *A_out, double *b_out);

/* -------------------------------------------------------------- solves -- */
/* coordinateDescent!(x, f, ProxL1(lambda0[, omega]), options)
 * (coordinate_descent.jl:7-39).  omega == NULL is ProxL1{T,Nothing}.
 * x in/out as the SparseIterate triple; with options->warmStart == 0 the input
 * iterate is ignored (fill!(x,0), :25) and the internal lambda continuation of
 * :28-36 runs.  stats may be NULL. */
int cdgpu_solve(cdgpu_handle h, double lambda0, const double *omega, const cdgpu_options *opt, double *nzval,
                int64_t *nzval2ind, int64_t *nnz, cdgpu_stats *stats);

/* Warm-started lambda path: for i in 1..m: coordinateDescent!(x, f,
 * ProxL1(lambda[i], omega), opt); beta_path[i] = copy(x); stop after the first i
 * with nnz(x) > max_hat_s (lasso.jl:250-257; max_hat_s < 0 == Inf).  Starts from
 * x = 0.  Output is CSC over the m columns: colptr[m+1] (0-based offsets),
 * rowval (1-based), nzval, caller-allocated with `capacity` entries; *m_done is
 * the number of columns produced.  stats: array of m or NULL. */
int cdgpu_path(cdgpu_handle h, const double *lambda, int64_t m, const double *omega, const cdgpu_options *opt,
               int64_t max_hat_s, int64_t capacity, int64_t *colptr, int64_t *rowval, double *nzval,
               int64_t *m_done, cdgpu_stats *stats);

/* scaledLasso!(x, X, y, lambda, omega, IterLassoOptions) (lasso.jl:107-144) on an
 * LS handle; the sigma loop runs on the device.  *sigma_out = std(r) as returned
 * in LassoSolution (:143); stats->sigma = the sigma of the last ProxL1. */
int cdgpu_scaled_solve(cdgpu_handle h, double lambda, const double *omega, const cdgpu_iter_options *opt,
                       double *nzval, int64_t *nzval2ind, int64_t *nnz, double *sigma_out, cdgpu_stats *stats);

/* f.r (naive handles, length n) or f.Ax (QUAD, length p) after the last solve;
 * LassoSolution.residuals aliases f.r (lasso.jl:37). */
int cdgpu_state(cdgpu_handle h, double *out);
/* _stdX!(out, X) / _stdX!(out, w, X) (utils.jl:127-151) on a naive handle; w NULL
 * uses the unweighted form even on a WLS handle.  On a QUAD handle (w must be
 * NULL) it returns sqrt(A_jj), which is _stdX!(X) when A = X'X/n. */
int cdgpu_stdx(cdgpu_handle h, const double *w, double *out);
/* _findLambdaMax at x = 0 (coordinate_descent.jl:118-149) */
int cdgpu_lambda_max(cdgpu_handle h, const double *omega, double *out);

/* locpolyl1(X, z, y, zgrid, degree, kernel, lambda0, refit=false, options)
 * (varying_coefficient_lasso.jl:30-79): one kernel-weighted local-polynomial
 * lasso per grid point, all grid points solved concurrently on the device (one
 * CTA per local problem).  out is dense ep x m column-major, ep = p*(degree+1),
 * column index (j-1)(degree+1)+1+l as _expand_X! (:550-569).  The reference
 * warm-starts each grid point from its predecessor (:56,:68); the batch starts
 * every problem from 0 — solutions agree at convergence.  stats: array of m or
 * NULL.  Only grid points [m_begin, m_end) are solved (sharding hook). */
int cdgpu_vc_solve(const double *X, int64_t n, int64_t p, int64_t ldx, const double *z, const double *y,
                   const double *zgrid, int64_t m, int64_t m_begin, int64_t m_end, int degree, int kernel_kind,
                   double bandwidth, double lambda0, const cdgpu_options *opt, int device, double *out,
                   cdgpu_stats *stats);
/* locpolyl1(...; refit=true) (:71-76): as cdgpu_vc_solve, and outR[S, g] = (Xs' W Xs) \ (Xs' W y) on the expanded
 * coordinates S of every group with a non-zero coefficient (get_nonzero_coordinates!, :488-512), zero elsewhere.
 * On the device the normal equations come from the same moment blocks as the lasso (no pass over the data) and
 * are solved by one warp per grid point (Cholesky).  outR: dense ep x m like out. */
int cdgpu_vc_solve_refit(const double *X, int64_t n, int64_t p, int64_t ldx, const double *z, const double *y,
                         const double *zgrid, int64_t m, int64_t m_begin, int64_t m_end, int degree, int kernel_kind,
                         double bandwidth, double lambda0, const cdgpu_options *opt, int device, double *out,
                         double *outR, cdgpu_stats *stats);
/* locpolyl1 with the reference's warm-start chain (varying_coefficient_lasso.jl:56,68: `beta` is created once and every
 * grid point's coordinateDescent! starts from its predecessor's solution, values and list order).  The grid points of
 * [m_begin, m_end) are cut into runs of `chain` consecutive points; a run is solved by one warp, point after point,
 * each from the previous one's iterate; the first point of a run starts from zero.  chain = m_end - m_begin is the
 * reference's loop exactly (one sequential chain: same iterates, passes and visits as the CPU path); chain = 1 is
 * cdgpu_vc_solve / cdgpu_vc_solve_refit.  outR may be NULL (no refit).  CDGPU_ECAP when p*(degree+1) > 512. */
int cdgpu_vc_solve_chain(const double *X, int64_t n, int64_t p, int64_t ldx, const double *z, const double *y,
                         const double *zgrid, int64_t m, int64_t m_begin, int64_t m_end, int degree, int kernel_kind,
                         double bandwidth, double lambda0, const cdgpu_options *opt, int64_t chain, int device,
                         double *out, double *outR, cdgpu_stats *stats);

/* locpolyl1's results as the reference returns them: SparseMatrixCSC (varying_coefficient_lasso.jl:46-47 `spzeros(ep, m)`,
 * :69 `out[:, i] = beta`, :76 `outR[...]`).  Same solve as cdgpu_vc_solve_chain (chain = 1: every grid point from zero); the
 * stored entries (value != 0) are compacted on the device and only colptr and those entries cross the ABI instead of the
 * dense ep x m matrix.  colptr[m+1] 0-based offsets, rowval 1-based rows in increasing order, nzval; columns outside
 * [m_begin, m_end) are empty.  rowval / nzval (and rowvalR / nzvalR) are caller-allocated with `capacity` entries each;
 * CDGPU_ECAP when a matrix has more, with colptr[m] (colptrR[m]) = the entries it needs.  colptrR NULL: no refit. */
int cdgpu_vc_solve_csc(const double *X, int64_t n, int64_t p, int64_t ldx, const double *z, const double *y,
                       const double *zgrid, int64_t m, int64_t m_begin, int64_t m_end, int degree, int kernel_kind,
                       double bandwidth, double lambda0, const cdgpu_options *opt, int64_t chain, int device,
                       int64_t capacity, int64_t *colptr, int64_t *rowval, double *nzval, int64_t *colptrR,
                       int64_t *rowvalR, double *nzvalR, cdgpu_stats *stats);

/* refitLassoPath(path, X, Y) (lasso.jl:208-225), one support at a time: the least-squares coefficients on the
 * columns `support` (1-based indices, ns of them) of the handle's design — X[:, S] \ y for a naive-form handle
 * (w-weighted for CDWeightedLSLoss), A[S,S] \ (-b[S]) for a covariance-form handle.  Normal equations and Cholesky
 * on the device; CDGPU_EARG ("SingularException") when the selected columns are linearly dependent. */
int cdgpu_refit(cdgpu_handle h, const int64_t *support, int64_t ns, double *coef_out);

/* lvocv_locpolyl1(X, z, y, degree, hArr, kernelType, lambda0, options) (varying_coefficient_lasso.jl:81-137):
 * leave-one-out choice of the bandwidth.  For every bandwidth hArr[ih] and observation i: the local problem at
 * z0 = z_i with w_i = 0, sigma initialised by screening (utils.jl:79-92), <= 10 rounds of CD at lambda0*sigma with
 * sigma re-estimated on the device between rounds (:112-124), refit on the selected groups and the squared error of
 * the prediction of y_i (:127-131).  All numH*n problems are independent units of one batch (one warp each); the
 * reference's warm start from the previous observation (:99) is cut.  sqerr[q], q = ih*n + i, is written for
 * q in [q_begin, q_end) (sharding hook); MSE[ih] = sum_i sqerr[ih*n + i].  stats: numH*n entries or NULL. */
int cdgpu_vc_lvocv(const double *X, int64_t n, int64_t p, int64_t ldx, const double *z, const double *y, int degree,
                   const double *hArr, int64_t numH, int kernel_kind, double lambda0, const cdgpu_options *opt,
                   int64_t q_begin, int64_t q_end, int device, double *sqerr, cdgpu_stats *stats);

#ifdef __cplusplus
}
#endif
#endif /* CDGPU_H */
