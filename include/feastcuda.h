/* libfeastcuda -- C ABI of the B200-native FEAST contour-integration engine.
 *
 * The reference (subhk/FeastKit.jl) is pure Julia and has no FFI boundary of its own; the drop-in
 * boundary is the set of Julia method signatures listed in SURVEY.md §8b.  Each entry point below is
 * what the Julia shim (feastkit.jl_b200/julia/FeastCUDA.jl) `ccall`s from the body of the
 * corresponding reference method; file:line citations are relative to /root/reference/src.
 *
 * Conventions
 *   - extern "C", plain pointers + sizes, every function returns an int status
 *     (FEASTCUDA_OK = 0; nonzero = failure, text via feastcuda_last_error()).  FEAST outcomes
 *     (non-convergence, M = 0, ...) are NOT statuses: they come back in *info with the reference's
 *     FeastError codes (core/feast_types.jl:257-268).
 *   - complex numbers are interleaved (re, im) doubles, exactly Julia's ComplexF64.
 *   - dense blocks are column-major with leading dimension n (Julia Matrix layout).
 *   - the caller owns every host array; the library copies inputs at set_* time and owns all
 *     device memory inside the handle; nothing calls back into the host runtime.
 *   - a handle is bound to one CUDA device and is not thread-safe; distinct handles are independent.
 *   - there is no CPU fallback: every entry point that computes fails with FEASTCUDA_ERR_CUDA when
 *     no usable GPU is present.
 */
#ifndef FEASTCUDA_H
#define FEASTCUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct feastcuda_handle_s* feastcuda_handle;

enum { FEASTCUDA_OK = 0, FEASTCUDA_ERR_ARG = 1, FEASTCUDA_ERR_CUDA = 2, FEASTCUDA_ERR_NCCL = 3,
       FEASTCUDA_ERR_UNSUPPORTED = 4, FEASTCUDA_ERR_STATE = 5 };

enum { FEASTCUDA_A = 0, FEASTCUDA_B = 1 };
enum { FEASTCUDA_CSR = 0, FEASTCUDA_CSC = 1 };                /* SparseMatrixCSC is CSC, 1-based */
enum { FEASTCUDA_SYM = 0, FEASTCUDA_HERM = 1, FEASTCUDA_GEN = 2 };
enum { FEASTCUDA_SOLVER_DIRECT = 0,      /* cached per-node LU (dense / banded operators)                        */
       FEASTCUDA_SOLVER_BICGSTAB = 1,    /* lock-step block BiCGStab, one complex solve per node                 */
       FEASTCUDA_SOLVER_MSLANCZOS = 2 }; /* multi-shift two-pass Lanczos: ONE real recurrence serves all nodes    */
                                         /* (standard real-symmetric problems, real basis, FILTER_TRUE; other     */
                                         /* problems fall back to BICGSTAB inside the GPU engine)                 */
enum { FEASTCUDA_FILTER_REFERENCE = 0,  /* complex half-contour sum 2*w*Y, dense/feast_dense.jl:231 */
       FEASTCUDA_FILTER_TRUE = 1 };     /* rho = Re g: real part for real-symmetric pencils          */
enum { FEASTCUDA_SHARD_NODES = 0,       /* block node partition, parallel/feast_mpi.jl:36-43         */
       FEASTCUDA_SHARD_COLUMNS = 1,     /* every rank: all nodes, a slice of the RHS columns         */
       FEASTCUDA_SHARD_BALANCED = 2 };  /* nodes x column-slices balanced by measured iterations     */

/* Solver keywords of the reference drivers (solver, solver_tol, solver_maxiter, solver_restart:
 * sparse/feast_sparse.jl:246-252) plus the engine's documented extras. */
typedef struct {
  int32_t solver;        /* FEASTCUDA_SOLVER_*                                                        */
  double  tol;           /* solver_tol; 0 -> 10^-fpm[3] (sparse/feast_sparse.jl:266)                  */
  int32_t maxiter;       /* solver_maxiter (iterations per node solve)                                */
  int32_t restart;       /* solver_restart: true-residual restarts allowed per node solve             */
  double  inner_rel;     /* >0: a column also stops at ||r|| <= inner_rel*||r0|| (inexact FEAST)       */
  int32_t ritz_guess;    /* 1: start node solves from X_j/(z - theta_j) of the previous loop          */
  int32_t filter;        /* FEASTCUDA_FILTER_*                                                        */
  int32_t shard;         /* FEASTCUDA_SHARD_* (only with feastcuda_nccl_init)                         */
  int32_t check_every;   /* host polls the device convergence flag every this many iterations (8)     */
  int32_t q0_real;       /* 1: Q0 is n x m0 REAL column-major (real-symmetric API)                    */
  int32_t x_real;        /* 1: X out is n x m0 REAL column-major = real.(q), dense/feast_dense.jl:372 */
  double  inner_rel0;    /* inner_rel used while no Ritz guess exists yet (loop 0); 0 -> inner_rel            */
  int32_t maxiter0;      /* iteration budget of those first solves; 0 -> maxiter                              */
  int32_t keep_going;    /* 1 (inexact mode only): M = 0 after a sweep does not abort, the sweep's Ritz      */
                         /*    vectors seed the next loop (the reference stops with info=5)                   */
  int32_t adaptive;      /* 1 (Lanczos path): once the eigen-residual is within reach of the tolerance, the sweep's     */
                         /*    inner target becomes 2*tol/epsout_prev (clamped to [1e-6, 0.1]) so that it is the last one */
  int32_t mixed;         /* 1, or 2 = when fpm[42] == 1 (Lanczos path, real symmetric): fpm[42] "single-precision solver" (core/feast_parameters.jl:316-319) --  */
                         /*    FP32 Krylov vectors and matrix entries, FP64 scalars/accumulator/Rayleigh-Ritz; falls back to    */
                         /*    FP64 vectors when a refined sweep gains less than a factor 4                                     */
  double  eps_floor;     /* > 0: the convergence tolerance is max(10^-fpm[3], eps_floor) -- the Float32 entry points      */
                         /*    (sfeast_*, cfeast_*) pass sqrt(eps(Float32)), core/feast_parameters.jl:398-405              */
  double  b_delta;       /* generalized Lanczos path (B Hermitian positive definite): accuracy of the Chebyshev inner solves  */
                         /*    with B in the first sweep (0 -> 1e-4); later sweeps use min(b_delta, 1e-3 * inner target)       */
} feastcuda_solver_opts;

typedef struct {
  int64_t loops, node_solves, krylov_iters, col_iters, spmm_launches, kernel_launches;
  int64_t ortho_passes, jacobi_sweeps, allreduce_bytes;
  double  ms_total, ms_h2d, ms_d2h, ms_solve, ms_ortho, ms_project, ms_eig, ms_resid, ms_allreduce;
  double  ms_spmm_sampled;   /* CUDA-event time of the sampled SpMM launches                          */
  int64_t spmm_sampled;      /* how many launches were sampled                                        */
  double  bytes_spmm_alg;    /* algorithmic bytes of ONE shifted SpMM launch at the last shape        */
  int64_t node_iters[128];   /* BiCGStab iterations of the last refinement loop, per node             */
  /* multi-shift Lanczos */
  int64_t lz_steps_p1, lz_steps_p2;   /* Lanczos steps run in pass 1 / pass 2 (all loops)                       */
  double  ms_lz_p1, ms_lz_p2;         /* wall time of the two passes                                            */
  /* CUDA-event samples per kernel kind (FEASTCUDA_KERN_*): total ms, launches sampled, algorithmic bytes of  */
  /* ONE launch at the last sampled shape                                                                      */
  double  ms_kern[8];
  int64_t n_kern[8];
  double  bytes_kern[8];
  double  ms_dev_run;                 /* CUDA-event time (library stream) of the feastcuda_run_interval calls    */
  int64_t lz_steps_fp32;              /* pass-1 Lanczos steps run with FP32 vectors (opts.mixed)                 */
  int64_t cheb_degree;                /* generalized Lanczos path: Chebyshev steps per inner solve with B (last sweep) */
} feastcuda_stats;

enum { FEASTCUDA_KERN_SPMM_Z = 0,   /* complex shifted SpMM (BiCGStab)            */
       FEASTCUDA_KERN_LZ_P1 = 1,    /* Lanczos pass-1 SpMM + dot                   */
       FEASTCUDA_KERN_LZ_UPD = 2,   /* Lanczos pass-1 vector update + norm         */
       FEASTCUDA_KERN_LZ_P2 = 3,    /* Lanczos pass-2 fused SpMM + accumulate      */
       FEASTCUDA_KERN_LZ_RES = 4,   /* Ritz-residual start block                   */
       FEASTCUDA_KERN_LZ_CHEB = 5,  /* one Chebyshev step of the inner solve with B (generalized Lanczos filter) */
       FEASTCUDA_KERN_BAND_LU = 6,  /* band LU of every node that needs a factor (one launch)                     */
       FEASTCUDA_KERN_BAND_SOLVE = 7 };/* band forward/backward substitution of all nodes x columns (one launch)  */

/* ---- lifetime --------------------------------------------------------------------------------- */
int feastcuda_create(feastcuda_handle* h, int device);
int feastcuda_destroy(feastcuda_handle h);
const char* feastcuda_last_error(feastcuda_handle h);   /* h may be NULL: last error of create()  */
int feastcuda_version(void);

/* ---- parameters and contours (host-side, tiny) -------------------------------------------------
 * feastinit!/feastdefault! core/feast_parameters.jl:7-18,41-386; returns FEASTCUDA_ERR_ARG where the
 * reference throws ArgumentError. fpm has 64 entries, addressed fpm[k-1] for the reference's fpm[k]. */
int feastcuda_feastinit(int64_t* fpm);
int feastcuda_feastdefault(int64_t* fpm);
/* feast_contour core/feast_tools.jl:212-284: half ellipse, ne = fpm[2] nodes; Zne/Wne: 2*ne doubles */
int feastcuda_contour(double Emin, double Emax, int64_t* fpm, double* Zne, double* Wne);
/* feast_gcontour core/feast_tools.jl:286-371: full rotated ellipse, ne = fpm[8] nodes */
int feastcuda_gcontour(double Emid_re, double Emid_im, double r, int64_t* fpm, double* Zne, double* Wne);

/* ---- operators ---------------------------------------------------------------------------------
 * Sparse: SparseMatrixCSC{Tv,Int} passes colptr/rowval/nzval with fmt=CSC, index_base=1.  CSC of a
 * symmetric matrix is its CSR; of a Hermitian matrix the conjugate (conjugated on load); a general
 * matrix is transposed once at set time.  `which` = FEASTCUDA_A or FEASTCUDA_B; B unset = identity. */
int feastcuda_set_csr_d(feastcuda_handle h, int which, int64_t n, int64_t nnz, const int64_t* ptr,
                        const int64_t* idx, const double* val, int index_base, int fmt, int structure);
int feastcuda_set_csr_z(feastcuda_handle h, int which, int64_t n, int64_t nnz, const int64_t* ptr,
                        const int64_t* idx, const double* val /* 2*nnz */, int index_base, int fmt, int structure);
int feastcuda_clear_b(feastcuda_handle h);
/* Matrix-free REAL SYMMETRIC standard problems: feast_matvec(A_mul!, B_mul!, N, interval) interfaces/feast_interfaces.jl:465-481
 * -> feast_sparse_matvec! sparse/feast_sparse.jl:1284-1471 (there: host closures + per-column GMRES).  Here the operator is a DEVICE
 * callback: `apply(ctx, n, ncols, X, ldx, Y, ldy, stream)` must enqueue on `stream` (a cudaStream_t) work that writes
 * Y[i*ldy + c] = sum_j A[i,j] X[j*ldx + c] for 0 <= c < ncols (row-major blocks of doubles in device memory; X and Y never
 * alias; the same input must give the same bits on every call: pass 2 of the Lanczos filter replays pass 1).  The solve uses the
 * multi-shift Lanczos filter (B = I only); every other operator/solver combination returns FEASTCUDA_ERR_UNSUPPORTED. */
typedef void (*feastcuda_apply_fn)(void* ctx, int64_t n, int64_t ncols, const double* X, int64_t ldx, double* Y, int64_t ldy, void* stream);
int feastcuda_set_matfree_d(feastcuda_handle h, int64_t n, feastcuda_apply_fn apply_a, void* ctx);

/* Dense column-major n x n, lda >= n (feast_syev!/sygv!/heev!/hegv!/geev!/gegv! dense/feast_dense.jl) */
int feastcuda_set_dense_d(feastcuda_handle h, int which, int64_t n, const double* a, int64_t lda, int structure);
int feastcuda_set_dense_z(feastcuda_handle h, int which, int64_t n, const double* a, int64_t lda, int structure);
/* LAPACK band storage: SYM/HERM upper (k+1) x n, diagonal in row k (0-based) banded/feast_banded.jl:205-214;
 * GEN (2k+1) x n, diagonal in row k banded/feast_banded.jl:263-271 */
int feastcuda_set_band_d(feastcuda_handle h, int which, int64_t n, int64_t k, const double* ab, int64_t ldab, int structure);
int feastcuda_set_band_z(feastcuda_handle h, int which, int64_t n, int64_t k, const double* ab, int64_t ldab, int structure);

/* ---- the solve: _feast_sparse_hermitian sparse/feast_sparse.jl:246-499, _feast_dense_complex_hermitian
 * dense/feast_dense.jl:78-351, _feast_banded_complex_hermitian banded/feast_banded.jl:561-823 (via
 * feast_scsrev!/scsrgv!/hcsrev!/hcsrgv!, feast_syev!/sygv!/heev!/hegv!, feast_sbev!/hbev!/...).
 * Zne/Wne: the contour (ne nodes; custom-contour `x` variants pass theirs).  Q0: initial subspace
 * n x m0 (complex column-major, or real if opts->q0_real); NULL -> library seed.  Outputs sized for m0:
 * lambda[m0], X n x m0, res[m0]; *M eigenpairs are valid (the shim trims). */
int feastcuda_solve_interval(feastcuda_handle h, double Emin, double Emax, int64_t m0, int64_t* fpm,
                             const double* Zne, const double* Wne, int64_t ne, const double* Q0,
                             const feastcuda_solver_opts* opts, double* lambda, double* X, double* res,
                             int64_t* M, int64_t* info, double* epsout, int64_t* loop);
/* Same solve split in three, so a caller can keep inputs resident in HBM:
 * upload Q0 -> run on device-resident data -> fetch results. */
int feastcuda_upload_subspace(feastcuda_handle h, int64_t m0, const double* Q0, int q0_real);
int feastcuda_run_interval(feastcuda_handle h, double Emin, double Emax, int64_t m0, int64_t* fpm,
                           const double* Zne, const double* Wne, int64_t ne,
                           const feastcuda_solver_opts* opts, int64_t* M, int64_t* info, double* epsout, int64_t* loop);
int feastcuda_fetch_results(feastcuda_handle h, int64_t m0, int x_real, double* lambda, double* X, double* res);

/* General (non-Hermitian) problems: feast_grci! kernel/feast_kernel.jl:646-962 driven as in
 * feast_gcsrgv! sparse/feast_sparse.jl:873-1006 / feast_gegv! dense/feast_dense.jl:402-593.
 * lambda: 2*m0 doubles (complex). */
int feastcuda_solve_contour(feastcuda_handle h, double Emid_re, double Emid_im, double r, int64_t m0, int64_t* fpm,
                            const double* Zne, const double* Wne, int64_t ne, const double* Q0,
                            const feastcuda_solver_opts* opts, double* lambda, double* X, double* res,
                            int64_t* M, int64_t* info, double* epsout, int64_t* loop);

/* ---- stage-level entry points (host buffers; used by the parity tests, the RCI shim and the
 * roofline bench).  X/Y/RHS/W/Q: n x m complex column-major. ------------------------------------- */
/* Y = z*(B X) - A X : SparseShiftedOperator mul! sparse/feast_sparse.jl:20-27,142-148 */
int feastcuda_spmm_shifted(feastcuda_handle h, double z_re, double z_im, int64_t m, const double* X, double* Y);
/* Y = A X (which=A) or B X (which=B): mul!(aq,A,q) sparse/feast_sparse.jl:392,402 */
int feastcuda_apply(feastcuda_handle h, int which, int64_t m, const double* X, double* Y);
/* (zB - A) Xout = RHS, all m columns in lock step: solve_shifted_iterative! sparse/feast_sparse.jl:164-236
 * (direct: ldiv! of the cached LU, dense/feast_dense.jl:196-207, banded gbtrf/gbtrs banded/feast_banded.jl:108,141).
 * X0 may be NULL.  iters[m], resid[m] (true residual norms) out. */
int feastcuda_block_solve(feastcuda_handle h, double z_re, double z_im, int64_t m, const double* RHS, const double* X0,
                          const feastcuda_solver_opts* opts, double* Xout, int64_t* iters, double* resid);
/* Qacc += w * Y : dense/feast_dense.jl:231, kernel/feast_kernel.jl:762-766 */
int feastcuda_accumulate(feastcuda_handle h, double w_re, double w_im, int64_t m, const double* Y, double* Qacc);
/* _feast_qr_compress! core/feast_aux.jl:101-131: orthonormal basis of the numerical range of W[:,1:m] */
int feastcuda_orthonormalize(feastcuda_handle h, int64_t n, int64_t m, const double* W, double rank_tol, double* Q, int64_t* rank);
/* C = X^H Y (m x m, column-major complex): mul!(zSq, adjoint(q), aq) dense/feast_dense.jl:253 */
int feastcuda_gram(feastcuda_handle h, int64_t n, int64_t m, const double* X, const double* Y, double* C);
/* eigen(Hermitian(Sq), Hermitian(Aq)) dense/feast_dense.jl:272 (Aq NULL = identity); column-major r x r in,
 * ascending lambda[r] and Aq-orthonormal V (column-major) out */
int feastcuda_reduced_eig(feastcuda_handle h, int64_t r, const double* Sq, const double* Aq, double* lambda, double* V, int64_t* sweeps);
/* Y (n x b) = X (n x a) * T (a x b), all complex column-major: the back-projection q = Q_proj * V of the RCI kernels
 * (kernel/feast_kernel.jl:187,547,838-845) and of the drivers (dense/feast_dense.jl:287-290) */
int feastcuda_rowtransform(feastcuda_handle h, int64_t n, int64_t a, int64_t b, const double* X, const double* T, double* Y);
/* eigen(A, B) of the small general reduced pencil (r <= 128; B NULL = identity): kernel/feast_kernel.jl:175,539,812
 * (LAPACK zggev in the reference).  A, B column-major complex r x r; lambda: 2*r doubles; V: r x r column-major, unit 2-norm columns.
 * A rank-deficient B (FEAST moment matrices when M0 exceeds the eigenvalue count inside) is deflated by a column-pivoted QR:
 * the remaining directions come back as lambda = +inf with the coordinate vector of a pivoted-out column (V stays regular), the finite pairs from the leading block */
int feastcuda_eig_general(feastcuda_handle h, int64_t r, const double* A, const double* B, double* lambda, double* V);
/* res_j = ||A x_j - lambda_j B x_j|| / max(|lambda_j|,1): dense/feast_dense.jl:309-322 */
int feastcuda_residuals(feastcuda_handle h, int64_t m, const double* X, const double* lambda /* complex, 2*m */, double* res);

/* ---- multi-GPU: one process per GPU; quadrature-node work is sharded and the n x m0 accumulator is
 * summed with ONE ncclAllReduce per refinement loop (MPI.Allreduce of Q_proj, parallel/feast_mpi.jl:
 * 119,341,858,1001).  The 128-byte id is created on rank 0 and distributed by the host
 * (torch.distributed / MPI / files). */
int feastcuda_nccl_unique_id(char* id128);
int feastcuda_nccl_init(feastcuda_handle h, int nranks, int rank, const char* id128);
/* Row sharding (after feastcuda_nccl_init, before the operator is uploaded; real symmetric sparse standard problems on the
 * multi-shift Lanczos filter): every rank owns a contiguous block of rows of A and of every block vector (the block rule of
 * parallel/feast_mpi.jl:36-43 applied to rows), halo rows are read from the owner's HBM over NVLink inside the gather kernels, the
 * per-step dot products are one-shot reductions through peer memory, Gram matrices and residual norms are summed by small
 * all-reduces; Q0 / X stay GLOBAL n x m0 host arrays of which each rank reads / writes its own rows only. */
int feastcuda_set_row_sharding(feastcuda_handle h, int on);
/* rows [row0, row0 + nrows) of the global order nglobal are this rank's (nrows = nglobal without row sharding) */
int feastcuda_row_range(feastcuda_handle h, int64_t* row0, int64_t* nrows, int64_t* nglobal);
/* node block owned by `rank`: (start, count), 0-based -- parallel/feast_mpi.jl:36-43 */
int feastcuda_node_partition(int64_t ne, int nranks, int rank, int64_t* start, int64_t* count);

int feastcuda_get_stats(feastcuda_handle h, feastcuda_stats* out);
int feastcuda_reset_stats(feastcuda_handle h);

#ifdef __cplusplus
}
#endif
#endif /* FEASTCUDA_H */
