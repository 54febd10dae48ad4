"""Small host-side helpers of the reference's API surface (no device work of their own).

feast_set_defaults! (interfaces/feast_interfaces.jl:483-511), eigvals_feast / eigen_feast (:421-445), feast_custom_contour
(:514-540) with feast_customcontour (core/feast_tools.jl:378-398), the rational-filter evaluators feast_rational(x) /
feast_grational(x) (core/feast_tools.jl:483-615), feast_summary (:542-561), feast_validate_interval with its Gershgorin bounds
(:563-640) and feast_memory_estimate (core/feast_aux.jl:645-664; here for the DEVICE workspace of the engine as well).
"""
from __future__ import annotations

import sys

import numpy as np

__all__ = [
    "feast_set_defaults", "eigvals_feast", "eigen_feast", "feast_customcontour", "feast_custom_contour", "feast_rationalx",
    "feast_rational", "feast_grationalx", "feast_grational", "feast_rational_expert", "feast_summary", "feast_validate_interval",
    "feast_memory_estimate",
]


def feast_set_defaults(fpm, print_level=1, integration_points=8, tolerance_exp=12, max_refinement=20):
    """feast_set_defaults!(fpm; print_level, integration_points, tolerance_exp, max_refinement)."""
    if len(fpm) < 64:
        raise ValueError("fpm array must have at least 64 elements")
    if print_level > 1:
        raise ValueError(f"print_level must be 0, 1, or negative for file output, got {print_level}")
    if integration_points <= 0:
        raise ValueError(f"integration_points must be positive, got {integration_points}")
    if not 0 <= tolerance_exp <= 16:
        raise ValueError(f"tolerance_exp must be between 0 and 16, got {tolerance_exp}")
    if max_refinement <= 0:
        raise ValueError(f"max_refinement must be positive, got {max_refinement}")
    fpm[0], fpm[1], fpm[2], fpm[3] = int(print_level), int(integration_points), int(tolerance_exp), int(max_refinement)
    return fpm


def eigvals_feast(A, *args, **kw):
    """eigvals_feast(A[, B], interval; kwargs...) -> eigenvalues only."""
    from . import feast
    return feast(A, *args, **kw).lambda_


def eigen_feast(A, *args, **kw):
    """eigen_feast(A[, B], interval; kwargs...) -> (values, vectors) like LinearAlgebra.Eigen."""
    from . import feast
    r = feast(A, *args, **kw)
    return r.lambda_, r.q


def feast_customcontour(Zne, fpm):
    """feast_customcontour(Zne, fpm): trapezoidal weights W_i = (Z_{i+1} - Z_{i-1}) / (2 ne) of a closed polygon of nodes;
    sets fpm[2] = ne."""
    Z = np.asarray(Zne, dtype=np.complex128)
    ne = len(Z)
    if ne == 0:
        raise ValueError("custom contour needs at least one node")
    fpm[1] = ne
    W = (np.roll(Z, -1) - np.roll(Z, 1)) / (2 * ne)
    return Z.copy(), W


def feast_custom_contour(nodes, A, *args, M0=10, fpm=None, **kw):
    """feast_custom_contour(nodes, A[, B], interval; M0, fpm): FEAST on the caller's nodes with trapezoidal weights."""
    import scipy.sparse as sp
    from . import feastinit
    from .api import (feast_hcsrevx, feast_hcsrgvx, feast_heevx, feast_hegvx, feast_scsrevx, feast_scsrgvx, feast_syevx,
                      feast_sygvx, ishermitian, issymmetric)
    if len(args) == 1:
        B, interval = None, args[0]
    elif len(args) == 2:
        B, interval = args
    else:
        raise TypeError("feast_custom_contour(nodes, A, interval) or feast_custom_contour(nodes, A, B, interval)")
    fpm = feastinit() if fpm is None else fpm
    Z, W = feast_customcontour(nodes, fpm)
    Emin, Emax = interval
    M0 = min(int(M0), A.shape[0])
    cplx = np.iscomplexobj(A) or (B is not None and np.iscomplexobj(B))
    if cplx and not ishermitian(A):
        raise ValueError("Matrix must be Hermitian for the interval solver; use feast_general")
    if not cplx and not issymmetric(A):
        raise ValueError("Matrix must be symmetric for the interval solver; use feast_general")
    if sp.issparse(A):
        if B is None:
            return (feast_hcsrevx if cplx else feast_scsrevx)(A, Emin, Emax, M0, fpm, Z, W, **kw)
        return (feast_hcsrgvx if cplx else feast_scsrgvx)(A, sp.csc_matrix(B), Emin, Emax, M0, fpm, Z, W, **kw)
    if B is None:
        return (feast_heevx if cplx else feast_syevx)(np.asarray(A), Emin, Emax, M0, fpm, Z, W, **kw)
    return (feast_hegvx if cplx else feast_sygvx)(np.asarray(A), np.asarray(B), Emin, Emax, M0, fpm, Z, W, **kw)


# ---- rational filter values (what the quadrature does to an eigenvalue) ----------------------------------------------------
def _zw(Zne, Wne):
    Z, W = np.asarray(Zne, dtype=np.complex128), np.asarray(Wne, dtype=np.complex128)
    if Z.shape != W.shape:
        raise ValueError("Zne and Wne must have the same length")
    return Z, W


def feast_rationalx(Zne, Wne, lambda_):
    """f(x) = 2 Re sum_e Wne[e] / (Zne[e] - x) for real x (half contour).  Also accepts the reference's convenience order
    feast_rationalx(lambda, Zne, Wne)."""
    if np.isrealobj(np.asarray(Zne)) and np.iscomplexobj(np.asarray(Wne)):
        Zne, Wne, lambda_ = Wne, lambda_, Zne
    Z, W = _zw(Zne, Wne)
    lam = np.asarray(lambda_, dtype=np.float64)
    return 2.0 * np.real((W[None, :] / (Z[None, :] - lam[:, None])).sum(axis=1))


feast_rational_expert = feast_rationalx


def feast_rational(lambda_, Emin, Emax, fpm):
    """feast_rational(lambda, Emin, Emax, fpm): filter values on the default half-ellipse contour (Zolotarev: unsupported)."""
    from . import feast_contour
    Z, W = feast_contour(float(Emin), float(Emax), fpm)
    return feast_rationalx(Z, W, lambda_)


def feast_grationalx(Zne, Wne, lambda_):
    """f(x) = sum_e Wne[e] / (Zne[e] - x) for complex x (full contour, no factor 2)."""
    if np.asarray(Wne).shape != np.asarray(Zne).shape:          # convenience order (lambda, Zne, Wne)
        Zne, Wne, lambda_ = Wne, lambda_, Zne
    Z, W = _zw(Zne, Wne)
    lam = np.asarray(lambda_, dtype=np.complex128)
    return (W[None, :] / (Z[None, :] - lam[:, None])).sum(axis=1)


def feast_grational(lambda_, Emid, r, fpm):
    """feast_grational(lambda, Emid, r, fpm): filter values on the default full-ellipse contour."""
    from . import feast_gcontour
    Z, W = feast_gcontour(complex(Emid), float(r), fpm)
    return feast_grationalx(Z, W, lambda_)


# ---- result / input analysis --------------------------------------------------------------------------------------------------
def feast_summary(result, io=None):
    """feast_summary([io,] result): the reference's text summary."""
    io = sys.stdout if io is None else io
    print("FeastKit Eigenvalue Solution Summary", file=io)
    print("=" * 40, file=io)
    print("Eigenvalues found: ", result.M, file=io)
    print("Final residual: ", result.epsout, file=io)
    print("Refinement loops: ", result.loop, file=io)
    print("Exit status: ", "Success" if result.info == 0 else f"Error {result.info}", file=io)
    if result.M > 0:
        print("\nEigenvalues:", file=io)
        for i in range(result.M):
            print(f"  λ[{i + 1}] = ", result.lambda_[i], "  (residual: ", result.res[i], ")", file=io)


def _gershgorin_bounds(A):
    import scipy.sparse as sp
    if sp.issparse(A):
        A = sp.csr_matrix(A)
        diag = np.real(A.diagonal())
        radii = np.asarray(abs(A).sum(axis=1)).ravel() - np.abs(A.diagonal())
    else:
        A = np.asarray(A)
        diag = np.real(np.diag(A))
        radii = np.abs(A).sum(axis=1) - np.abs(np.diag(A))
    return float((diag - radii).min()), float((diag + radii).max())


def feast_validate_interval(A, interval):
    """feast_validate_interval(A, (Emin, Emax)) -> Gershgorin estimate (min, max) of the spectrum; warns when the interval lies
    outside it, ArgumentError (ValueError) for Emin >= Emax."""
    import warnings
    Emin, Emax = interval
    if Emin >= Emax:
        raise ValueError("Invalid interval: Emin must be less than Emax")
    lo, hi = _gershgorin_bounds(A)
    if Emax < lo or Emin > hi:
        warnings.warn(f"Search interval [{Emin}, {Emax}] may not contain eigenvalues. Estimated eigenvalue range: [{lo}, {hi}]")
    return lo, hi


def feast_memory_estimate(N, M0, precision=np.float64, device=False, io=None):
    """feast_memory_estimate(N, M0, T): bytes of the reference's host workspace (core/feast_aux.jl:645-664).  device=True: bytes of
    the ENGINE's HBM workspace instead (12 complex n x M0 block slots)."""
    io = sys.stdout if io is None else io
    sz = np.dtype(precision).itemsize
    if device:
        total = 12 * N * M0 * 16
        print(f"libfeastcuda device workspace: {total / 1024 ** 2:.2f} MB (12 block slots of n x M0 complex128)", file=io)
        return total
    work, workc = N * M0 * sz, N * M0 * 2 * sz
    reduced, eigen = 2 * M0 * M0 * sz, (N * M0 + 2 * M0) * sz
    total = work + workc + reduced + eigen
    print("FeastKit Memory Estimate:", file=io)
    for name, v in (("Workspace (real)", work), ("Workspace (complex)", workc), ("Reduced matrices", reduced), ("Eigendata", eigen),
                    ("Total estimate", total)):
        print(f"  {name}: {v / 1024 ** 2:.2f} MB", file=io)
    return total
