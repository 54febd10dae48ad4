"""Reference-named drivers (the bodies a FeastKit.jl maintainer would replace with ``ccall``s).

Names, positional arguments and keyword arguments follow the reference; ``name!`` is spelled ``name``.
Engine extras are keyword-only and documented in DESIGN.md: ``Q0`` (initial subspace; Julia's seeded RNG
stream is not reproducible outside Julia), ``inner_rel`` / ``ritz_guess`` (inexact inner solves),
``filter`` ("reference" = the reference's complex half-contour sum, "true" = its real part), ``shard``.
"""
from __future__ import annotations

import numpy as np

from . import _lib as L
from ._host import column_major, dense_equals_own_adjoint

__all__ = [
    "feast_scsrev", "feast_scsrgv", "feast_hcsrev", "feast_hcsrgv", "feast_scsrevx", "feast_scsrgvx", "feast_hcsrevx",
    "feast_hcsrgvx", "feast_syev", "feast_sygv", "feast_heev", "feast_hegv", "feast_syevx", "feast_sygvx", "feast_heevx",
    "feast_hegvx", "feast_sbev", "feast_sbgv", "feast_hbev", "feast_hbgv", "feast", "feast_banded", "issymmetric",
    "ishermitian", "feast_matvec", "feast_sparse_matvec", "feast_general", "feast_gcsrev", "feast_gcsrgv", "feast_geev", "feast_gegv", "feast_gbev", "feast_gbgv",
]


def _eng(engine):
    from . import default_engine
    return engine if engine is not None else default_engine()


def issymmetric(M):
    import scipy.sparse as sp
    if sp.issparse(M):
        return (abs(M - M.T)).nnz == 0 if M.shape[0] == M.shape[1] else False
    M = np.asarray(M)
    return M.ndim == 2 and M.shape[0] == M.shape[1] and dense_equals_own_adjoint(M, False)


def ishermitian(M):
    import scipy.sparse as sp
    if sp.issparse(M):
        return (abs(M - M.conj().T)).nnz == 0 if M.shape[0] == M.shape[1] else False
    M = np.asarray(M)
    return M.ndim == 2 and M.shape[0] == M.shape[1] and dense_equals_own_adjoint(M, True)


M0_CAP = 128


def _check_m0_cap(M0):
    """Known divergence from the reference (which accepts any 0 < M0 <= N): the device block kernels hold at most 128 columns."""
    if int(M0) > M0_CAP:
        raise ValueError(f"M0 = {int(M0)} exceeds the {M0_CAP}-column cap of the GPU block kernels: slice the interval into "
                         f"sub-intervals holding fewer eigenvalues (INTEGRATION.md, 'Known divergences')")


def _solver_kw(kind, solver, solver_tol, solver_maxiter, solver_restart, extras):
    """Map the reference keywords (sparse/feast_sparse.jl:246-266) onto feastcuda_solver_opts."""
    solver_choice = "gmres" if solver == "iterative" else solver
    if solver_choice not in ("direct", "gmres", "bicgstab", "mslanczos"):
        raise ValueError(f"Unsupported solver option '{solver}'. Use :direct, :gmres, or :iterative.")
    kw = dict(solver_tol=solver_tol, solver_maxiter=solver_maxiter)
    # sparse inputs are served by Krylov solvers (there is no sparse LU in the engine): the multi-shift Lanczos
    # recurrence where it applies (real symmetric, B = I, real basis -- the engine falls back to the lock-step
    # block BiCGStab otherwise); solver="bicgstab" forces the per-node complex solves.
    # solver_restart keeps its meaning of "restart budget" (true-residual restarts of the short recurrence)
    if kind != "sparse":
        kw["solver"] = "direct" if solver_choice == "direct" else "bicgstab"
    else:
        kw["solver"] = "bicgstab" if solver_choice == "bicgstab" else "mslanczos"
    kw["solver_restart"] = 3 if solver_restart is None else int(solver_restart)     # None = the caller left the keyword alone
    if kind == "sparse":
        # engine defaults for Krylov node solves (DESIGN.md "Inner solves"): start from the previous loop's Ritz pairs
        # and stop three digits below that guess's residual; inner_rel=0, ritz_guess=False gives the reference's
        # zero-guess / solve-to-tol / fail-with-info-5 behaviour (sparse/feast_sparse.jl:164-236,359-363)
        kw["ritz_guess"] = True
        kw["inner_rel"] = 1e-3
        kw["adaptive"] = True
        # fpm[42] ("mixed precision: 1 = single-precision solver", default 1, core/feast_parameters.jl:316-319): the engine honours it where
        # it has an FP32 inner solver (real symmetric standard problems on the Lanczos filter); mixed=False keeps FP64 Krylov vectors
        kw["mixed"] = "fpm"
    kw["filter"] = "true"
    for k in ("inner_rel", "ritz_guess", "filter", "shard", "check_every", "inner_rel0", "maxiter0", "keep_going", "adaptive", "eps_floor", "mixed", "b_delta"):
        if k in extras:
            kw[k] = extras.pop(k)
    if extras:
        raise TypeError(f"unexpected keyword arguments: {sorted(extras)}")
    return kw


def _hermitian_solve(kind, setA, setB, N, Emin, Emax, M0, fpm, real_result, contour=None, solver="direct",
                     solver_tol=0.0, solver_maxiter=500, solver_restart=None, Q0=None, engine=None, **extras):
    from . import check_feast_srci_input, feast_contour, feastdefault_
    feastdefault_(fpm)
    check_feast_srci_input(N, M0, float(Emin), float(Emax), fpm)
    _check_m0_cap(M0)
    kw = _solver_kw(kind, solver, solver_tol, solver_maxiter, solver_restart, extras)
    eng = _eng(engine)   # argument errors above are raised before any device is touched, as in the reference
    setA(eng)
    if setB is not None:
        setB(eng)
    else:
        eng.clear_b()
    eng.init_distributed()
    Zne, Wne = contour if contour is not None else feast_contour(float(Emin), float(Emax), fpm)
    return eng.solve_interval(float(Emin), float(Emax), int(M0), fpm, Zne, Wne, Q0=Q0, x_real=real_result, **kw)


# ---- sparse (sparse/feast_sparse.jl) ---------------------------------------------------------------------
def _sparse_pair(A, B, structure):
    import scipy.sparse as sp
    if not sp.issparse(A):
        raise TypeError("A must be a sparse matrix (SparseMatrixCSC)")
    N = A.shape[0]
    if A.shape[1] != N:
        raise ValueError("A must be square")
    if B is not None and B.shape != (N, N):
        raise ValueError("B must be same size as A")
    setA = lambda e: e.set_sparse(L.A, A, structure)
    setB = None if B is None else (lambda e: e.set_sparse(L.B, B, structure))
    return N, setA, setB


def feast_scsrev(A, Emin, Emax, M0, fpm, **kw):
    """feast_scsrev!(A, Emin, Emax, M0, fpm) -- sparse/feast_sparse.jl:1516-1529 (real symmetric, B = I)."""
    N, sA, _ = _sparse_pair(A.astype(np.float64), None, L.SYM)
    return _hermitian_solve("sparse", sA, None, N, Emin, Emax, M0, fpm, True, **kw)


def feast_scsrgv(A, B, Emin, Emax, M0, fpm, **kw):
    """feast_scsrgv!(A, B, Emin, Emax, M0, fpm; solver, ...) -- sparse/feast_sparse.jl:713-731."""
    N, sA, sB = _sparse_pair(A.astype(np.float64), B.astype(np.float64), L.SYM)
    return _hermitian_solve("sparse", sA, sB, N, Emin, Emax, M0, fpm, True, **kw)


def feast_hcsrev(A, Emin, Emax, M0, fpm, **kw):
    """feast_hcsrev! -- sparse/feast_sparse.jl:759-788 (complex Hermitian, B = I)."""
    if not ishermitian(A):
        raise ValueError("Matrix A must be Hermitian")
    N, sA, _ = _sparse_pair(A.astype(np.complex128), None, L.HERM)
    return _hermitian_solve("sparse", sA, None, N, Emin, Emax, M0, fpm, False, **kw)


def feast_hcsrgv(A, B, Emin, Emax, M0, fpm, **kw):
    """feast_hcsrgv! -- sparse/feast_sparse.jl:815-831."""
    if not ishermitian(A):
        raise ValueError("Matrix A must be Hermitian")
    if not ishermitian(B):
        raise ValueError("Matrix B must be Hermitian")
    N, sA, sB = _sparse_pair(A.astype(np.complex128), B.astype(np.complex128), L.HERM)
    return _hermitian_solve("sparse", sA, sB, N, Emin, Emax, M0, fpm, False, **kw)


def feast_scsrevx(A, Emin, Emax, M0, fpm, Zne, Wne, **kw):
    """Custom-contour form, sparse/feast_sparse.jl:751-757."""
    return feast_scsrev(A, Emin, Emax, M0, fpm, contour=(Zne, Wne), **kw)


def feast_scsrgvx(A, B, Emin, Emax, M0, fpm, Zne, Wne, **kw):
    return feast_scsrgv(A, B, Emin, Emax, M0, fpm, contour=(Zne, Wne), **kw)


def feast_hcsrevx(A, Emin, Emax, M0, fpm, Zne, Wne, **kw):
    return feast_hcsrev(A, Emin, Emax, M0, fpm, contour=(Zne, Wne), **kw)


def feast_hcsrgvx(A, B, Emin, Emax, M0, fpm, Zne, Wne, **kw):
    return feast_hcsrgv(A, B, Emin, Emax, M0, fpm, contour=(Zne, Wne), **kw)


# ---- dense (dense/feast_dense.jl) ------------------------------------------------------------------------
def _dense_pair(A, B, structure, dtype):
    A = np.asarray(A, dtype=dtype)
    N = A.shape[0]
    if A.ndim != 2 or A.shape[1] != N:
        raise ValueError("Matrix A must be square")
    if B is not None:
        B = np.asarray(B, dtype=dtype)
        if B.shape != (N, N):
            raise ValueError("Matrix B must match size of A")
    if not ishermitian(A):
        raise ValueError("Matrix A must be Hermitian")
    if B is not None and not ishermitian(B):
        raise ValueError("Matrix B must be Hermitian positive definite")
    setA = lambda e: e.set_dense(L.A, A, structure)
    setB = None if B is None else (lambda e: e.set_dense(L.B, B, structure))
    return N, setA, setB


def feast_syev(A, Emin, Emax, M0, fpm, **kw):
    """feast_syev! -- dense/feast_dense.jl:776-797."""
    N, sA, _ = _dense_pair(A, None, L.SYM, np.float64)
    return _hermitian_solve("dense", sA, None, N, Emin, Emax, M0, fpm, True, **kw)


def feast_sygv(A, B, Emin, Emax, M0, fpm, **kw):
    """feast_sygv! -- dense/feast_dense.jl:356-370."""
    N, sA, sB = _dense_pair(A, B, L.SYM, np.float64)
    return _hermitian_solve("dense", sA, sB, N, Emin, Emax, M0, fpm, True, **kw)


def feast_heev(A, Emin, Emax, M0, fpm, **kw):
    """feast_heev! -- dense/feast_dense.jl:390-400."""
    N, sA, _ = _dense_pair(A, None, L.HERM, np.complex128)
    return _hermitian_solve("dense", sA, None, N, Emin, Emax, M0, fpm, False, **kw)


def feast_hegv(A, B, Emin, Emax, M0, fpm, **kw):
    """feast_hegv! -- dense/feast_dense.jl:799-810."""
    N, sA, sB = _dense_pair(A, B, L.HERM, np.complex128)
    return _hermitian_solve("dense", sA, sB, N, Emin, Emax, M0, fpm, False, **kw)


def feast_syevx(A, Emin, Emax, M0, fpm, Zne, Wne, **kw):
    return feast_syev(A, Emin, Emax, M0, fpm, contour=(Zne, Wne), **kw)


def feast_sygvx(A, B, Emin, Emax, M0, fpm, Zne, Wne, **kw):
    return feast_sygv(A, B, Emin, Emax, M0, fpm, contour=(Zne, Wne), **kw)


def feast_heevx(A, Emin, Emax, M0, fpm, Zne, Wne, **kw):
    return feast_heev(A, Emin, Emax, M0, fpm, contour=(Zne, Wne), **kw)


def feast_hegvx(A, B, Emin, Emax, M0, fpm, Zne, Wne, **kw):
    return feast_hegv(A, B, Emin, Emax, M0, fpm, contour=(Zne, Wne), **kw)


# ---- banded (banded/feast_banded.jl) ---------------------------------------------------------------------
def _band_pair(A, ka, B, kb, structure, dtype):
    A = np.asarray(A, dtype=dtype)
    N = A.shape[1]
    if A.shape[0] < ka + 1:
        raise ValueError("A matrix storage insufficient for kla")
    if B is not None:
        B = np.asarray(B, dtype=dtype)
        if B.shape[0] < kb + 1:
            raise ValueError("B matrix storage insufficient for klb")
        if B.shape[1] != N:
            raise ValueError("B must be same size as A")
    setA = lambda e: e.set_band(L.A, A, ka, structure)
    setB = None if B is None else (lambda e: e.set_band(L.B, B, kb, structure))
    return N, setA, setB


def _smom_banded(N, setA, setB, Emin, Emax, M0, fpm, Q0=None, engine=None, contour=None):
    """The reference's own route for real symmetric bands (banded/feast_banded.jl:9-186): the moment kernel feast_srci! driven
    with one band LU per node (FACTORIZE), block solves of (Ze B - A) X = B work (SOLVE) and band mat-vecs (MULT_A), every piece
    on the device (feastcuda_block_solve / feastcuda_apply and the stage calls inside feast_srci)."""
    from . import FeastResult, check_feast_srci_input, feastdefault_
    from .rci import FeastRCIState, Ref, feast_srci
    feastdefault_(fpm)
    check_feast_srci_input(N, M0, float(Emin), float(Emax), fpm)
    eng = _eng(engine)
    setA(eng)
    if setB is not None:
        setB(eng)
    else:
        eng.clear_b()
    ijob, Ze, eps, loop, mode, info = Ref(-1), Ref(0j), Ref(0.0), Ref(0), Ref(0), Ref(-1)
    work, workc = np.zeros((N, M0)), np.zeros((N, M0), dtype=np.complex128)
    Aq, Sq = np.zeros((M0, M0)), np.zeros((M0, M0))
    lam, q, res = np.zeros(M0), np.zeros((N, M0)), np.zeros(M0)
    if Q0 is not None:                                   # fpm[5] = 1: the caller's initial subspace travels in `work`
        fpm[4] = 1
        work[:, :M0] = np.asarray(Q0, dtype=np.float64)
    state, z = FeastRCIState(), 0j
    ne = len(contour[0]) if contour is not None else int(fpm[1])
    for _ in range((int(fpm[3]) + 2) * (2 * ne + 2) + 8):      # every refinement loop issues 2 jobs per node + MULT_A; +1 for INIT
        feast_srci(ijob, N, Ze, work, workc, Aq, Sq, fpm, eps, loop, float(Emin), float(Emax), M0, lam, q, mode, res, info, state=state,
                   engine=eng, contour=contour)
        if ijob.v == 0 or (ijob.v == -1 and info.v != 0):      # DONE, or INIT refused the arguments (info = FeastError code)
            break
        if ijob.v == 10:
            z = Ze.v
        elif ijob.v == 11:
            rhs = work[:, :M0].astype(np.complex128)
            if setB is not None:
                rhs = eng.apply(L.B, rhs)
            workc[:, :M0] = eng.block_solve(z, rhs, solver="direct")[0]
        elif ijob.v == 30:
            work[:, :mode.v] = eng.apply(L.A, np.asarray(q[:, :mode.v], dtype=np.complex128)).real
    M = int(state.M)
    return FeastResult(lam[:M].copy(), q[:, :M].copy(), M, res[:M].copy(), int(info.v), float(eps.v), int(loop.v), eng.stats())




def feast_sbev(A, kla, Emin, Emax, M0, fpm, method="hrr", **kw):
    """feast_sbev! -- banded/feast_banded.jl:1410-1420 (real symmetric band, upper storage).  method="smom": the reference's
    moment route through feast_srci! (see feast_sbgv)."""
    N, sA, _ = _band_pair(A, kla, None, 0, L.SYM, np.float64)
    if method == "smom":
        return _smom_banded(N, sA, None, Emin, Emax, M0, fpm, Q0=kw.pop("Q0", None), engine=kw.pop("engine", None), contour=kw.pop("contour", None))
    return _hermitian_solve("band", sA, None, N, Emin, Emax, M0, fpm, True, **kw)


def feast_sbgv(A, B, kla, klb, Emin, Emax, M0, fpm, method="hrr", **kw):
    """feast_sbgv! -- banded/feast_banded.jl:9-186.  The reference drives the moment kernel feast_srci! with band LUs here;
    method="hrr" (default) runs the engine's QR-compress + Rayleigh-Ritz loop on the same band kernels (same converged pairs,
    residuals with B), method="smom" reproduces the reference's route step for step: feast_srci! on the stage calls, band LU
    solves and band mat-vecs on the device, residuals ||A q - lambda q|| without B as kernel/feast_kernel.jl:250 has them."""
    N, sA, sB = _band_pair(A, kla, B, klb, L.SYM, np.float64)
    if method == "smom":
        return _smom_banded(N, sA, sB, Emin, Emax, M0, fpm, Q0=kw.pop("Q0", None), engine=kw.pop("engine", None), contour=kw.pop("contour", None))
    return _hermitian_solve("band", sA, sB, N, Emin, Emax, M0, fpm, True, **kw)


def feast_hbev(A, kla, Emin, Emax, M0, fpm, **kw):
    """feast_hbev! -- banded/feast_banded.jl:326-383."""
    N, sA, _ = _band_pair(A, kla, None, 0, L.HERM, np.complex128)
    return _hermitian_solve("band", sA, None, N, Emin, Emax, M0, fpm, False, **kw)


def feast_hbgv(A, B, kla, klb, Emin, Emax, M0, fpm, **kw):
    """feast_hbgv! -- banded/feast_banded.jl:385-421."""
    N, sA, sB = _band_pair(A, kla, B, klb, L.HERM, np.complex128)
    return _hermitian_solve("band", sA, sB, N, Emin, Emax, M0, fpm, False, **kw)


# ---- general (non-Hermitian) problems: kernel/feast_kernel.jl:646-962 behind the storage drivers ---------------------
def _general_solve(kind, setA, setB, N, Emid, r, M0, fpm, contour=None, solver="direct", solver_tol=0.0, solver_maxiter=None,
                   solver_restart=None, Q0=None, engine=None, **extras):
    from . import check_feast_grci_input, feast_gcontour, feastdefault_
    feastdefault_(fpm)
    check_feast_grci_input(N, M0, complex(Emid), float(r), fpm)
    solver_choice = "gmres" if solver == "iterative" else solver
    if solver_choice not in ("direct", "gmres", "bicgstab"):
        raise ValueError(f"Unsupported solver option '{solver}'. Use :direct, :gmres, or :iterative.")
    _check_m0_cap(M0)
    # the reference's defaults (500 iterations, restart 30) belong to its per-column GMRES; the engine's lock-step Krylov solves get their
    # own defaults only when the caller left the keywords alone -- explicit values are passed through unchanged
    if solver_maxiter is None:
        solver_maxiter = 2000 if kind == "sparse" else 500
    if solver_restart is None:
        solver_restart = 3
    # sparse pencils: ONE two-sided Lanczos recurrence per column serves every node of the contour (the engine keeps the per-node block
    # BiCGStab when B is not diagonally dominant or the block is wider than 64 columns per rank); solver="bicgstab" forces the latter
    kw = dict(solver_tol=solver_tol, solver_maxiter=int(solver_maxiter),
              solver=("bicgstab" if solver_choice == "bicgstab" else "mslanczos") if kind == "sparse" else "direct",
              solver_restart=int(solver_restart))
    if kind == "sparse":
        kw.update(ritz_guess=True, inner_rel=1e-3, adaptive=True)
    for k in ("inner_rel", "shard", "check_every", "eps_floor", "ritz_guess", "adaptive", "inner_rel0", "maxiter0", "b_delta"):
        if k in extras:
            kw[k] = extras.pop(k)
    if extras:
        raise TypeError(f"unexpected keyword arguments: {sorted(extras)}")
    eng = _eng(engine)
    setA(eng)
    if setB is not None:
        setB(eng)
    else:
        eng.clear_b()
    eng.init_distributed()
    Zne, Wne = contour if contour is not None else feast_gcontour(complex(Emid), float(r), fpm)
    return eng.solve_contour(complex(Emid), float(r), int(M0), fpm, Zne, Wne, Q0=Q0, **kw)


def feast_gcsrgv(A, B, Emid, r, M0, fpm, **kw):
    """feast_gcsrgv!(A, B, Emid, r, M0, fpm; solver, ...) -- sparse/feast_sparse.jl:873-1006."""
    N, sA, sB = _sparse_pair(A.astype(np.complex128), None if B is None else B.astype(np.complex128), L.GEN)
    return _general_solve("sparse", sA, sB, N, Emid, r, M0, fpm, **kw)


def feast_gcsrev(A, Emid, r, M0, fpm, **kw):
    """feast_gcsrev! -- sparse/feast_sparse.jl:1531-1551 (B = I)."""
    return feast_gcsrgv(A, None, Emid, r, M0, fpm, **kw)


def _dense_general(A, B):
    A = np.asarray(A, dtype=np.complex128)
    N = A.shape[0]
    if A.ndim != 2 or A.shape[1] != N:
        raise ValueError("Matrix A must be square")
    if B is not None:
        B = np.asarray(B, dtype=np.complex128)
        if B.shape != (N, N):
            raise ValueError("Matrix B must match size of A")
    setA = lambda e: e.set_dense(L.A, A, L.GEN)
    setB = None if B is None else (lambda e: e.set_dense(L.B, B, L.GEN))
    return N, setA, setB


def feast_gegv(A, B, Emid, r, M0, fpm, **kw):
    """feast_gegv! -- dense/feast_dense.jl:402-593."""
    N, sA, sB = _dense_general(A, B)
    return _general_solve("dense", sA, sB, N, Emid, r, M0, fpm, **kw)


def feast_geev(A, Emid, r, M0, fpm, **kw):
    """feast_geev! -- dense/feast_dense.jl:812-828."""
    return feast_gegv(A, None, Emid, r, M0, fpm, **kw)


def feast_gbgv(A, B, kla, klb, Emid, r, M0, fpm, **kw):
    """feast_gbgv!: general band (2k+1) x n storage -- banded/feast_banded.jl:1548-1574, 1088-1284."""
    A = np.asarray(A, dtype=np.complex128)
    N = A.shape[1]
    if A.shape[0] < 2 * kla + 1:
        raise ValueError("A matrix storage insufficient for kla")
    setA = lambda e: e.set_band(L.A, A, kla, L.GEN)
    setB = None
    if B is not None:
        B = np.asarray(B, dtype=np.complex128)
        if B.shape[0] < 2 * klb + 1 or B.shape[1] != N:
            raise ValueError("B matrix storage insufficient for klb")
        setB = lambda e: e.set_band(L.B, B, klb, L.GEN)
    return _general_solve("band", setA, setB, N, Emid, r, M0, fpm, **kw)


def feast_gbev(A, kla, Emid, r, M0, fpm, **kw):
    """feast_gbev! -- banded/feast_banded.jl:1576-1600."""
    return feast_gbgv(A, None, kla, 0, Emid, r, M0, fpm, **kw)


def feast_general(A, *args, M0=10, fpm=None, **kw):
    """feast_general(A, center, radius; M0, fpm) / feast_general(A, B, center, radius; ...) -- interfaces/feast_interfaces.jl:274-379."""
    import scipy.sparse as sp
    from . import feastinit
    _select_backend(kw)
    if len(args) == 2:
        B, (center, radius) = None, args
    elif len(args) == 3:
        B, center, radius = args
    else:
        raise TypeError("feast_general(A, center, radius) or feast_general(A, B, center, radius)")
    if A.shape[0] != A.shape[1]:
        raise ValueError("Matrix must be square")
    fpm = feastinit() if fpm is None else fpm
    M0 = min(int(M0), A.shape[0])
    if sp.issparse(A):
        return feast_gcsrgv(sp.csc_matrix(A), None if B is None else sp.csc_matrix(B), center, radius, M0, fpm, **kw)
    return feast_gegv(np.asarray(A), None if B is None else np.asarray(B), center, radius, M0, fpm, **kw)


# ---- backend keywords of the high-level API (interfaces/feast_interfaces.jl:24-58, test/test_backend_api.jl) -----------------
_BACKENDS = ("serial", "auto", "threads", "distributed", "mpi")


def _normalize_parallel(parallel):
    if parallel is True:
        return "auto"
    if parallel is False:
        return "serial"
    if isinstance(parallel, str):
        return parallel.lstrip(":")
    raise ValueError(f"Invalid parallel option: {parallel}")


def _normalize_backend(parallel, backend):
    """`parallel` is the legacy keyword, `backend` its explicit replacement; both only when they agree."""
    if backend is not None:
        requested = str(backend).lstrip(":")
        if parallel is not None:
            legacy = _normalize_parallel(parallel)
            if legacy != requested:
                raise ValueError(f"Conflicting backend requests: backend={requested} and parallel={legacy}")
    elif parallel is not None:
        requested = _normalize_parallel(parallel)
    else:
        requested = "serial"
    if requested not in _BACKENDS:
        raise ValueError(f"Unknown backend: {requested}. Use :serial, :auto, :threads, :distributed, or :mpi")
    return requested


def _select_backend(kw):
    """Pops backend/parallel/strict_backend/use_threads/comm from the keywords and validates them like the reference.  Every
    accepted choice runs on the GPU engine: :serial and :auto always; :distributed / :mpi mean "all ranks of the running
    torch.distributed job" and are refused (ArgumentError -> ValueError) when there is no such job, as the reference refuses
    them without workers / MPI; there is no threads backend (one process drives one GPU), so an explicit :threads is refused
    the way the reference refuses it for inputs its threaded path does not serve."""
    backend, parallel = kw.pop("backend", None), kw.pop("parallel", None)
    strict = bool(kw.pop("strict_backend", False))
    kw.pop("use_threads", None)
    kw.pop("comm", None)
    requested = _normalize_backend(parallel, backend)
    fallback = (not strict) and (backend in ("auto", ":auto") or (backend is None and (parallel is True or parallel in ("auto", ":auto"))))
    if requested in ("serial", "auto"):
        return requested
    if requested == "threads":
        if fallback:
            return "serial"
        raise ValueError("Backend :threads is not available: libfeastcuda drives one GPU per process (use :serial, :auto, or "
                         ":distributed/:mpi under torchrun)")
    world = 1
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            world = dist.get_world_size()
    except Exception:   # noqa: BLE001
        world = 1
    if world <= 1:
        if fallback:
            return "serial"
        raise ValueError(f"Backend :{requested} requested but no multi-process job is running (launch one process per GPU with torchrun)")
    return requested


# ---- high level (interfaces/feast_interfaces.jl:143-272, 381-420) ----------------------------------------
def feast(A, *args, M0=10, fpm=None, **kw):
    """feast(A, (Emin,Emax); M0, fpm) / feast(A, B, (Emin,Emax); M0, fpm) -- interfaces/feast_interfaces.jl:143-272."""
    import scipy.sparse as sp
    from . import feastinit
    _select_backend(kw)
    if len(args) == 1:
        B, interval = None, args[0]
    elif len(args) == 2:
        B, interval = args
    else:
        raise TypeError("feast(A, interval) or feast(A, B, interval)")
    Emin, Emax = interval
    N = A.shape[0]
    if A.shape[0] != A.shape[1]:
        raise ValueError("Matrix must be square")
    fpm = feastinit() if fpm is None else fpm
    M0 = min(int(M0), N)
    cplx = np.iscomplexobj(A) or (B is not None and np.iscomplexobj(B))
    if cplx:
        if not ishermitian(A):
            raise ValueError("Matrix must be Hermitian for the interval solver; use feast_general")
    elif not issymmetric(A):
        raise ValueError("Matrix must be symmetric for the interval solver; use feast_general")
    if sp.issparse(A):
        if cplx:
            return feast_hcsrev(A, Emin, Emax, M0, fpm, **kw) if B is None else feast_hcsrgv(A, sp.csc_matrix(B), Emin, Emax, M0, fpm, **kw)
        return feast_scsrev(A, Emin, Emax, M0, fpm, **kw) if B is None else feast_scsrgv(A, sp.csc_matrix(B), Emin, Emax, M0, fpm, **kw)
    if cplx:
        return feast_heev(A, Emin, Emax, M0, fpm, **kw) if B is None else feast_hegv(A, B, Emin, Emax, M0, fpm, **kw)
    return feast_syev(A, Emin, Emax, M0, fpm, **kw) if B is None else feast_sygv(A, B, Emin, Emax, M0, fpm, **kw)


def feast_banded(A, kla, interval, B=None, klb=0, M0=10, fpm=None, **kw):
    """feast_banded(A, kla, interval; B, klb, M0, fpm) -- interfaces/feast_interfaces.jl:381-420."""
    from . import feastinit
    _select_backend(kw)
    Emin, Emax = interval
    fpm = feastinit() if fpm is None else fpm
    A = np.array(A)
    M0 = min(int(M0), A.shape[1])
    if np.iscomplexobj(A):
        return feast_hbev(A, kla, Emin, Emax, M0, fpm, **kw) if B is None else feast_hbgv(A, np.array(B), kla, klb, Emin, Emax, M0, fpm, **kw)
    return feast_sbev(A, kla, Emin, Emax, M0, fpm, **kw) if B is None else feast_sbgv(A, np.array(B), kla, klb, Emin, Emax, M0, fpm, **kw)


# ---- matrix-free (interfaces/feast_interfaces.jl:465-481, sparse/feast_sparse.jl:1284-1471) -------------------------------
def _is_identity_operator(B_mul, N, engine):
    """The reference's feast_matvec always takes a B_mul!; the engine's matrix-free path is the standard problem, so a B
    operator is accepted iff it acts as the identity on a random device block."""
    import torch
    dev = torch.device("cuda", _eng(engine).device)
    X = torch.randn(N, 2, dtype=torch.float64, device=dev)
    Y = torch.zeros_like(X)
    B_mul(Y, X)
    return bool(torch.equal(Y, X))


def feast_sparse_matvec(A_matvec, B_matvec, N, Emin, Emax, M0, fpm, **kw):
    """feast_sparse_matvec!(A_matvec!, B_matvec!, N, Emin, Emax, M0, fpm) -- sparse/feast_sparse.jl:1284-1471.

    A_matvec is the DEVICE operator: a Python callable A_matvec(Y, X) on torch CUDA tensor views (n x ncols blocks; fill
    Y = A @ X), or a (function pointer, ctx) pair for a compiled feastcuda_apply_fn.  The reference runs one GMRES per (node,
    column) around host closures; here one multi-shift Lanczos recurrence per column calls the operator once per step for all
    columns.  B_matvec: None, or an operator that acts as the identity (anything else: NotImplementedError)."""
    if N <= 0:
        raise ValueError("Matrix size N must be positive")
    if B_matvec is not None and not _is_identity_operator(B_matvec, int(N), kw.get("engine")):
        raise NotImplementedError("matrix-free generalized problems (B != I) are not supported by the multi-shift Lanczos filter")
    if isinstance(A_matvec, tuple):
        setA = lambda e: e.set_matfree(int(N), A_matvec[0], A_matvec[1])
    else:
        setA = lambda e: e.set_matfree(int(N), A_matvec)
    return _hermitian_solve("sparse", setA, None, int(N), Emin, Emax, M0, fpm, True, **kw)


def feast_matvec(A_mul, B_mul, N, interval, M0=10, fpm=None, **kw):
    """feast_matvec(A_mul!, B_mul!, N, interval; M0, fpm) -- interfaces/feast_interfaces.jl:465-481."""
    from . import feastinit
    Emin, Emax = interval
    fpm = feastinit() if fpm is None else fpm
    return feast_sparse_matvec(A_mul, B_mul, N, Emin, Emax, M0, fpm, **kw)


# ---- precision / parallel alias families (interfaces/feast_precision_aliases.jl:10-971) ------------------
# d/z = Float64/ComplexF64 names forward 1:1; pd/pz = "parallel" aliases: with comm/use_threads the reference
# picks an MPI/threads backend, here every visible rank of the torch.distributed job takes part.
def _alias(fn):
    def wrapper(*a, comm=None, use_threads=None, **kw):
        return fn(*a, **kw)
    wrapper.__doc__ = f"alias of {fn.__name__} (interfaces/feast_precision_aliases.jl)"
    return wrapper


_ALIASES = {
    "dfeast_scsrev": feast_scsrev, "dfeast_scsrgv": feast_scsrgv, "zfeast_hcsrev": feast_hcsrev, "zfeast_hcsrgv": feast_hcsrgv,
    "dfeast_syev": feast_syev, "dfeast_sygv": feast_sygv, "zfeast_heev": feast_heev, "zfeast_hegv": feast_hegv,
    "dfeast_sbev": feast_sbev, "dfeast_sbgv": feast_sbgv, "zfeast_hbev": feast_hbev, "zfeast_hbgv": feast_hbgv,
    "dfeast_scsrevx": feast_scsrevx, "dfeast_scsrgvx": feast_scsrgvx, "zfeast_hcsrevx": feast_hcsrevx, "zfeast_hcsrgvx": feast_hcsrgvx,
    "dfeast_syevx": feast_syevx, "dfeast_sygvx": feast_sygvx, "zfeast_heevx": feast_heevx, "zfeast_hegvx": feast_hegvx,
    "zfeast_gcsrev": feast_gcsrev, "zfeast_gcsrgv": feast_gcsrgv, "zfeast_geev": feast_geev, "zfeast_gegv": feast_gegv,
    "zfeast_gbev": feast_gbev, "zfeast_gbgv": feast_gbgv, "zifeast_gcsrev": feast_gcsrev, "zifeast_gcsrgv": feast_gcsrgv,
}
for _name, _fn in list(_ALIASES.items()):
    globals()[_name] = _alias(_fn)
    globals()["p" + _name] = _alias(_fn)
    __all__ += [_name, "p" + _name]


# Float32 / ComplexF32 families (interfaces/feast_precision_aliases.jl:10-117, 163-423; ps/pc: 497-771): inputs are widened,
# the engine computes in Float64 and stops at the reference's single-precision tolerance max(10^-fpm[3], sqrt(eps(Float32)))
# (core/feast_parameters.jl:398-405); results are narrowed to Float32 / ComplexF32 like FeastResult{Float32,...}.
_EPS32 = float(np.sqrt(np.finfo(np.float32).eps))


def _alias32(fn):
    def wrapper(*a, comm=None, use_threads=None, **kw):
        from . import FeastResult
        args = [x.astype(np.complex128 if np.iscomplexobj(x) else np.float64) if hasattr(x, "astype") and hasattr(x, "shape") else x
                for x in a]
        if kw.get("Q0") is not None:
            kw["Q0"] = np.asarray(kw["Q0"]).astype(np.complex128 if np.iscomplexobj(kw["Q0"]) else np.float64)
        if fn in (feast_scsrev, feast_scsrevx):
            kw.setdefault("mixed", True)      # single-precision names: FP32 Krylov vectors (the outer loop stays FP64)
        r = fn(*args, eps_floor=_EPS32, **kw)
        cq = np.complex64 if np.iscomplexobj(r.q) else np.float32
        cl = np.complex64 if np.iscomplexobj(r.lambda_) else np.float32
        return FeastResult(r.lambda_.astype(cl), r.q.astype(cq), r.M, r.res.astype(np.float32), r.info, float(np.float32(r.epsout)),
                           r.loop, r.stats)
    wrapper.__doc__ = f"single-precision alias of {fn.__name__} (interfaces/feast_precision_aliases.jl)"
    return wrapper


for _name, _fn in list(_ALIASES.items()):
    _n32 = {"d": "s", "z": "c"}[_name[0]] + _name[1:]
    globals()[_n32] = _alias32(_fn)
    globals()["p" + _n32] = _alias32(_fn)
    __all__ += [_n32, "p" + _n32]
