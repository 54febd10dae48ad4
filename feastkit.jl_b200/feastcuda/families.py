"""Complex-symmetric and polynomial driver families -- host-side names over the general-contour engine.

Complex symmetric (A == transpose(A), complex): `feast_geev_complex_sym!/feast_gegv_complex_sym!` (dense/feast_dense.jl:1261-1286),
`feast_scsrev_complex!/feast_scsrgv_complex!` (sparse/feast_sparse.jl:1038-1095), `feast_sbev_complex!/feast_sbgv_complex!`
(banded/feast_banded.jl:1469-1525) and their custom-contour `x` forms.  The reference projects with the transpose-bilinear form
Q^T A Q (CS-RR); the engine's one-sided Rayleigh-Ritz on the orthonormalised filtered block (`feastcuda_solve_contour`) finds the
same eigenvalues inside the contour with right eigenvectors normalised to unit 2-norm, so these names validate the symmetry
exactly like the reference (ArgumentError -> ValueError, same message) and forward to the general drivers.

Polynomial problems P(lambda) q = 0, P = A[0] + lambda A[1] + ... + lambda^d A[d]: `feast_pep!` (dense/feast_dense.jl:715-772)
builds the first companion linearisation of size d*N, solves it with `feast_gegv!` using M0*d columns and returns the first N
components of the eigenvectors; `feast_gepev!/feast_hepev!/feast_sypev!` (+x) and `feast_polynomial`
(interfaces/feast_interfaces.jl:448-462) are wrappers.
"""
from __future__ import annotations

import numpy as np

from .api import feast_gbgv, feast_gcsrgv, feast_gegv

__all__ = [
    "feast_geev_complex_sym", "feast_gegv_complex_sym", "feast_scsrev_complex", "feast_scsrgv_complex", "feast_scsrevx_complex",
    "feast_scsrgvx_complex", "feast_sbev_complex", "feast_sbgv_complex", "feast_sbevx_complex", "feast_sbgvx_complex",
    "feast_pep", "feast_gepev", "feast_gepevx", "feast_hepev", "feast_hepevx", "feast_sypev", "feast_sypevx", "feast_polynomial",
    "companion_linearization",
]


def check_complex_symmetric(M, sparse=False):
    """check_complex_symmetric (core/feast_aux.jl:665-668) / _check_complex_symmetric (sparse/feast_sparse.jl:93-95)."""
    if sparse:
        ok = abs(M - M.T).max() == 0 if M.nnz else True
        if not ok:
            raise ValueError("Matrix must be complex symmetric (equal to its transpose)")
    elif not np.array_equal(np.asarray(M), np.asarray(M).T):
        raise ValueError("Matrix must be complex-symmetric (equal to its transpose).")
    return True


# ---- dense ------------------------------------------------------------------------------------------------------------
def feast_gegv_complex_sym(A, B, Emid, r, M0, fpm, **kw):
    """feast_gegv_complex_sym!(A, B, Emid, r, M0, fpm; solver, ...) -- dense/feast_dense.jl:1274-1286, 1026-1259."""
    A = np.asarray(A, dtype=np.complex128)
    if A.ndim != 2 or A.shape[0] != A.shape[1]:
        raise ValueError("A must be square")
    if B is not None:
        B = np.asarray(B, dtype=np.complex128)
        if B.shape != A.shape:
            raise ValueError("B must be same size as A")
    check_complex_symmetric(A)
    if B is not None:
        check_complex_symmetric(B)
    return feast_gegv(A, B, Emid, r, M0, fpm, **kw)


def feast_geev_complex_sym(A, Emid, r, M0, fpm, **kw):
    """feast_geev_complex_sym! -- dense/feast_dense.jl:1261-1272."""
    return feast_gegv_complex_sym(A, None, Emid, r, M0, fpm, **kw)


# ---- sparse -----------------------------------------------------------------------------------------------------------
def feast_scsrgv_complex(A, B, Emid, r, M0, fpm, **kw):
    """feast_scsrgv_complex! -- sparse/feast_sparse.jl:1038-1051."""
    import scipy.sparse as sp
    A = sp.csc_matrix(A, dtype=np.complex128)
    check_complex_symmetric(A, sparse=True)
    if B is not None:
        B = sp.csc_matrix(B, dtype=np.complex128)
        check_complex_symmetric(B, sparse=True)
    return feast_gcsrgv(A, B, Emid, r, M0, fpm, **kw)


def feast_scsrev_complex(A, Emid, r, M0, fpm, **kw):
    """feast_scsrev_complex! -- sparse/feast_sparse.jl:1069-1080."""
    return feast_scsrgv_complex(A, None, Emid, r, M0, fpm, **kw)


def feast_scsrgvx_complex(A, B, Emid, r, M0, fpm, Zne, Wne, **kw):
    return feast_scsrgv_complex(A, B, Emid, r, M0, fpm, contour=(np.asarray(Zne, complex), np.asarray(Wne, complex)), **kw)


def feast_scsrevx_complex(A, Emid, r, M0, fpm, Zne, Wne, **kw):
    return feast_scsrgv_complex(A, None, Emid, r, M0, fpm, contour=(np.asarray(Zne, complex), np.asarray(Wne, complex)), **kw)


# ---- banded: symmetric upper band (k+1) x n, diagonal in the last row (banded/feast_banded.jl:423-440) ------------------
def _symmetric_band_to_general(AB, k):
    """(k+1) x n upper band of a complex SYMMETRIC matrix -> general band (2k+1) x n, diagonal in row k (no conjugation)."""
    AB = np.asarray(AB, dtype=np.complex128)
    if AB.shape[0] < k + 1:
        raise ValueError("A matrix storage insufficient for ka")
    n = AB.shape[1]
    GB = np.zeros((2 * k + 1, n), dtype=np.complex128)
    for d in range(k + 1):                       # d-th superdiagonal: A[j-d, j] = AB[k-d, j]
        GB[k - d, d:] = AB[k - d, d:]
        if d:
            GB[k + d, :n - d] = AB[k - d, d:]    # mirrored entry A[j, j-d]
    return GB


def feast_sbgv_complex(A, B, ka, kb, Emid, r, M0, fpm, **kw):
    """feast_sbgv_complex! -- banded/feast_banded.jl:1469-1481, 833-1078."""
    GA = _symmetric_band_to_general(A, int(ka))
    GB = None if B is None else _symmetric_band_to_general(B, int(kb))
    return feast_gbgv(GA, GB, int(ka), int(kb) if B is not None else 0, Emid, r, M0, fpm, **kw)


def feast_sbev_complex(A, ka, Emid, r, M0, fpm, **kw):
    """feast_sbev_complex! -- banded/feast_banded.jl:1499-1510."""
    return feast_sbgv_complex(A, None, ka, 0, Emid, r, M0, fpm, **kw)


def feast_sbgvx_complex(A, B, ka, kb, Emid, r, M0, fpm, Zne, Wne, **kw):
    return feast_sbgv_complex(A, B, ka, kb, Emid, r, M0, fpm, contour=(np.asarray(Zne, complex), np.asarray(Wne, complex)), **kw)


def feast_sbevx_complex(A, ka, Emid, r, M0, fpm, Zne, Wne, **kw):
    return feast_sbgv_complex(A, None, ka, 0, Emid, r, M0, fpm, contour=(np.asarray(Zne, complex), np.asarray(Wne, complex)), **kw)


# ---- polynomial eigenvalue problems -------------------------------------------------------------------------------------
def companion_linearization(coeffs):
    """First companion form of P(lambda) = sum_k lambda^k coeffs[k] (dense/feast_dense.jl:727-760): identity blocks on the block
    superdiagonal of A_lin, -coeffs[0..d-1] in its last block row; identity on the first d-1 diagonal blocks of B_lin and
    coeffs[d] in the last one.  A_lin y = lambda B_lin y with y = [q; lambda q; ...; lambda^(d-1) q]."""
    d = len(coeffs) - 1
    N = np.asarray(coeffs[0]).shape[0]
    DN = d * N
    A_lin = np.zeros((DN, DN), dtype=np.complex128)
    B_lin = np.zeros((DN, DN), dtype=np.complex128)
    eye = np.eye(N)
    for i in range(d - 1):
        A_lin[i * N:(i + 1) * N, (i + 1) * N:(i + 2) * N] = eye
        B_lin[i * N:(i + 1) * N, i * N:(i + 1) * N] = eye
    for j in range(d):
        A_lin[(d - 1) * N:, j * N:(j + 1) * N] = -np.asarray(coeffs[j])
    B_lin[(d - 1) * N:, (d - 1) * N:] = np.asarray(coeffs[d])
    return A_lin, B_lin


def feast_pep(A, d, Emid, r, M0, fpm, **kw):
    """feast_pep!(A, d, Emid, r, M0, fpm) -- dense/feast_dense.jl:715-772."""
    from . import FeastGeneralResult
    if len(A) != d + 1:
        raise ValueError("Need d+1 coefficient matrices")
    if d < 1:
        raise ValueError("Polynomial degree d must be at least 1")
    N = np.asarray(A[0]).shape[0]
    for Ai in A:
        if np.asarray(Ai).shape != (N, N):
            raise ValueError("All matrices must be same size")
    A_lin, B_lin = companion_linearization([np.asarray(Ai, dtype=np.complex128) for Ai in A])
    res = feast_gegv(A_lin, B_lin, complex(Emid), float(r), int(M0) * d, fpm, **kw)
    M = res.M
    return FeastGeneralResult(res.lambda_[:M], res.q[:N, :M], M, res.res[:M], res.info, res.epsout, res.loop, res.stats)


def feast_gepev(A, d, Emid, r, M0, fpm, **kw):
    """feast_gepev! -- dense/feast_dense.jl:946-949."""
    return feast_pep(A, d, Emid, r, M0, fpm, **kw)


def feast_hepev(A, d, Emid, r, M0, fpm, **kw):
    """feast_hepev! -- dense/feast_dense.jl:960-963."""
    return feast_pep(A, d, Emid, r, M0, fpm, **kw)


def feast_sypev(A, d, Emid, r, M0, fpm, **kw):
    """feast_sypev! (real coefficient matrices are widened) -- dense/feast_dense.jl:974-978."""
    return feast_pep([np.asarray(Ai, dtype=np.complex128) for Ai in A], d, Emid, r, M0, fpm, **kw)


def _with_contour(fn):
    def wrapper(A, d, Emid, r, M0, fpm, Zne, Wne, **kw):
        return fn(A, d, Emid, r, M0, fpm, contour=(np.asarray(Zne, complex), np.asarray(Wne, complex)), **kw)
    wrapper.__doc__ = f"custom-contour form of {fn.__name__} (with_custom_contour, dense/feast_dense.jl:951-987)"
    return wrapper


feast_gepevx = _with_contour(feast_gepev)
feast_hepevx = _with_contour(feast_hepev)
feast_sypevx = _with_contour(feast_sypev)


def feast_polynomial(coeffs, center, radius, M0=10, fpm=None, **kw):
    """feast_polynomial(coeffs, center, radius; M0, fpm) -- interfaces/feast_interfaces.jl:448-462."""
    from . import feastinit
    fpm = feastinit() if fpm is None else fpm
    return feast_pep(list(coeffs), len(coeffs) - 1, center, radius, M0, fpm, **kw)
