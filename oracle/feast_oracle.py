"""CPU oracle: a NumPy/SciPy restatement of FeastKit.jl's FEAST contour-integration path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it, and only as the checker or as the timed CPU arm.  The
product (``feastkit.jl_b200/``) never imports it and has no CPU fallback.

Pinning status
--------------
The reference is pure Julia and ``julia`` is absent from this image (and from
the GPU box), so the reference itself cannot be executed.  The oracle is pinned
against every known-answer case the reference's own test-suite holds for this
path (``tests/test_oracle_golden.py``): analytic
spectra / dense ``eigvals`` of the tiny matrices in ``test/runtests.jl``, the
exact helper outputs in ``test/test_allocation_helpers.jl:85-292`` and the
contour golden numbers.  The reference pins nothing at n >= 1000, no loop
counts and no eigenvectors beyond residuals: for the BASELINE.json sizes the
parity is "unpinned by the reference" and rests on analytic spectra of the
synthetic matrices plus this restatement at reduced n.

Third-party arithmetic the reference reaches that is not under /root/reference
(no Manifest.toml is committed; compat bounds from Project.toml:18-22):
LAPACK/OpenBLAS via Julia stdlib (zgetrf/zgetrs, zgeqp3, zhegv, zggev, zgbtrf/
zgbtrs), SuiteSparse UMFPACK, Krylov.jl >=0.10.1,<0.11 (gmres),
FastGaussQuadrature 1.x (gausslegendre).  Substitutes here: SciPy LAPACK
wrappers, SuperLU (``splu``), SciPy ``gmres``, ``numpy leggauss`` -- they solve
the same linear systems / reduced eigenproblems to tolerance, so converged
eigenpairs agree; bitwise iterates do not.

All ``file:line`` citations are relative to /root/reference/src.
``fpm`` is a length-64 integer list addressed with the reference's 1-based
index through ``fpm[k-1]``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp
import scipy.sparse.linalg as spla

FEAST_UNINITIALIZED = -111

# FeastError codes, core/feast_types.jl:257-268
SUCCESS, ERR_N, ERR_M0, ERR_EMIN_EMAX, ERR_EMID_R, ERR_NO_CONV, ERR_MEM, ERR_INTERNAL, ERR_LAPACK, ERR_FPM = range(10)
# FeastRCIJob, core/feast_types.jl:227-249
RCI_INIT, RCI_DONE, RCI_FACTORIZE, RCI_SOLVE, RCI_FACTORIZE_T, RCI_SOLVE_T, RCI_MULT_A, RCI_MULT_A_H, RCI_MULT_B, RCI_MULT_B_H = (
    -1, 0, 10, 11, 20, 21, 30, 31, 40, 41)


# --------------------------------------------------------------------------
# parameters: core/feast_parameters.jl
# --------------------------------------------------------------------------
def feastinit():
    """feastinit!: every entry = -111 sentinel (core/feast_parameters.jl:7-18)."""
    return [FEAST_UNINITIALIZED] * 64


def feastdefault(fpm):
    """feastdefault! (core/feast_parameters.jl:41-386).  Mutates and returns fpm.

    fpm[30] is never set by any reference caller, so the digit-conditional
    branches are dead; they are restated anyway for completeness.
    """
    if len(fpm) < 64:
        raise ValueError("fpm array must have at least 64 elements")
    U = FEAST_UNINITIALIZED
    g = lambda k: fpm[k - 1]

    def s(k, v):
        fpm[k - 1] = v

    dig = [0] * 7  # dig[1..6]
    if g(30) != U and g(30) > 0:
        rem = g(30)
        for i in range(1, 7):
            dig[7 - i] = rem % 10
            rem //= 10
    if g(1) == U:
        s(1, 0)
    elif g(1) > 1:
        raise ValueError("Invalid fpm[1]")
    if g(14) == U:
        s(14, 0)
    elif g(14) < 0 or g(14) > 2:
        raise ValueError("Invalid fpm[14]")
    if g(16) == U:
        s(16, 0)
        if dig[3] == 2:
            s(16, 1)
        if dig[4] == 3:
            s(16, 1)
        if dig[4] == 1 and dig[2] == 4:
            s(16, 1)
    elif g(16) < 0 or g(16) > 2:
        raise ValueError("Invalid fpm[16]")
    if g(16) == 2 and (dig[4] == 3 or (dig[4] == 1 and dig[2] == 4)):
        raise ValueError("Invalid fpm[16]=2")
    if g(2) == U or g(2) <= 0:
        s(2, 8)
        if dig[3] == 2:
            s(2, 4)
        if g(14) == 2:
            s(2, 3)
    elif g(16) in (0, 2) and g(2) > 20:
        if g(2) not in (24, 32, 40, 48, 56):
            raise ValueError("Invalid fpm[2]")
    if g(3) == U:
        s(3, 12)
    elif g(3) < 0 or g(3) > 16:
        raise ValueError("Invalid fpm[3]")
    if g(4) == U or g(4) <= 0:
        s(4, 20)
        if dig[3] == 2:
            s(4, 50)
    if g(5) == U:
        s(5, 0)
    elif g(5) not in (0, 1):
        raise ValueError("Invalid fpm[5]")
    if g(6) == U:
        s(6, 1)
    elif g(6) not in (0, 1):
        raise ValueError("Invalid fpm[6]")
    if g(7) == U:
        s(7, 5)
    elif g(7) < 0 or g(7) > 7:
        raise ValueError("Invalid fpm[7]")
    if g(8) == U or g(8) <= 0:
        s(8, 16)
        if dig[3] == 2:
            s(8, 8)
        if g(14) == 2:
            s(8, 6)
    elif g(8) < 2:
        raise ValueError("Invalid fpm[8]")
    elif g(16) == 0 and g(8) > 40:
        if g(8) not in (48, 64, 80, 96, 112):
            raise ValueError("Invalid fpm[8]")
    if g(9) == U:
        s(9, 0)
    if g(10) == U:
        s(10, 1)
        if dig[5] == 1:
            s(10, 0)
    elif g(10) not in (0, 1):
        raise ValueError("Invalid fpm[10]")
    for k in (11, 12):
        if g(k) == U:
            s(k, 0)
    if g(13) == U:
        s(13, 0)
    elif g(13) < 0 or g(13) > 3:
        raise ValueError("Invalid fpm[13]")
    if g(15) == U:
        s(15, 0)
        if dig[4] == 1:
            s(15, 2)
    elif g(15) < 0 or g(15) > 2:
        raise ValueError("Invalid fpm[15]")
    if g(14) == 2:
        s(15, 1)
    if g(17) == U:
        s(17, 0)
    if g(18) == U:
        s(18, 100)
        if dig[3] == 1 and dig[6] <= 5:
            if dig[4] == 2:
                s(18, 30)
            if dig[4] == 1 and dig[2] not in (3, 4):
                s(18, 30)
    elif g(18) < 0:
        raise ValueError("Invalid fpm[18]")
    if g(19) == U:
        s(19, 0)
    elif g(19) < -180 or g(19) > 180:
        raise ValueError("Invalid fpm[19]")
    for k in range(20, 29):
        if g(k) == U:
            s(k, 0)
    if g(29) == U:
        s(29, 0)
    if g(31) == U:
        s(31, 40)
    if g(32) == U:
        s(32, 10)
    for k in (33, 34, 35):
        if g(k) == U:
            s(k, 0)
    if g(36) == U:
        s(36, 1)
    if g(37) == U:
        s(37, 0)
    if g(38) == U:
        s(38, 1)
    for k in (39, 40):
        if g(k) == U:
            s(k, 0)
    if g(41) == U:
        s(41, 1)
    if g(42) == U:
        s(42, 1)
    for k in (43, 44):
        if g(k) == U:
            s(k, 0)
    if g(45) == U:
        s(45, 1)
    if g(46) == U:
        s(46, 40)
    for k in (47, 48, 49):
        if g(k) == U:
            s(k, 0)
    for k in range(50, 59):
        if g(k) == U:
            s(k, 0)
    for k in (59, 60, 61, 62, 63, 64):
        if g(k) == U:
            s(k, 0)
    return fpm


def feast_tolerance(fpm, dtype=np.float64):
    """feast_tolerance (core/feast_parameters.jl:391-405), incl. the Float32 floor sqrt(eps)."""
    e = fpm[2]
    tol = 1e-12 if (e < 0 or e > 16) else 10.0 ** (-e)
    if np.dtype(dtype) == np.float32:
        return max(float(np.float32(tol)), float(np.sqrt(np.finfo(np.float32).eps)))
    return tol


def check_feast_srci_input(N, M0, Emin, Emax, fpm):
    """core/feast_aux.jl:369-399 (throws ArgumentError in the reference)."""
    if N <= 0:
        raise ValueError("Matrix size N must be positive")
    if M0 <= 0 or M0 > N:
        raise ValueError("Number of eigenvalues M0 must be between 1 and N")
    if Emin >= Emax:
        raise ValueError("Search interval [Emin, Emax] must be valid")
    if len(fpm) < 64:
        raise ValueError("fpm array must have at least 64 elements")
    if 0 < fpm[1] < 3:
        raise ValueError("Number of integration points must be at least 3")
    return True


def check_feast_grci_input(N, M0, Emid, r, fpm):
    """core/feast_aux.jl:401-425."""
    if N <= 0:
        raise ValueError("Matrix size N must be positive")
    if M0 <= 0 or M0 > N:
        raise ValueError("Number of eigenvalues M0 must be between 1 and N")
    if r <= 0:
        raise ValueError("Contour radius must be positive")
    if len(fpm) < 64:
        raise ValueError("fpm array must have at least 64 elements")
    return True


# --------------------------------------------------------------------------
# contours: core/feast_tools.jl:212-398
# --------------------------------------------------------------------------
_ZOLO = None


def zolotarev_point(n, k):
    """zolotarev_point(n, k), core/feast_tools.jl:182-210: table value (tests/golden/zolotarev_tables.json, transcribed from the
    reference's ZOLOTAREV_TABLES by tests/golden/make_zolotarev.py) or the reference's fallback for an n without a table."""
    global _ZOLO
    if _ZOLO is None:
        import json
        import pathlib
        _ZOLO = json.loads((pathlib.Path(__file__).resolve().parents[1] / "tests" / "golden" / "zolotarev_tables.json").read_text())["tables"]
    t = _ZOLO.get(str(n))
    if t is not None:
        if k == 0:
            return 0j, complex(*t["we0"])
        if 1 <= k <= len(t["nodes"]):
            a, b, c, d = t["nodes"][k - 1]
            return complex(a, b), complex(c, d)
    if k == 0:
        return 0j, 1 + 0j
    theta = math.pi * (2 * k - 1) / (2 * n)
    return complex(math.cos(theta), math.sin(theta)), complex(0.0, math.pi / n)


def feast_contour(Emin, Emax, fpm):
    """Half-ellipse nodes/weights (core/feast_tools.jl:212-284)."""
    if fpm[1] == FEAST_UNINITIALIZED or fpm[1] <= 0:
        feastdefault(fpm)
    ne, fpm16, fpm18 = fpm[1], fpm[15], fpm[17]
    r = (Emax - Emin) / 2.0
    Emid = Emin + r
    aspect = fpm18 * 0.01
    ba, ab = -math.pi / 2, math.pi / 2
    Z = np.empty(ne, dtype=np.complex128)
    W = np.empty(ne, dtype=np.complex128)
    if fpm16 == 0:
        xg, wg = np.polynomial.legendre.leggauss(ne)  # ascending nodes, same as FastGaussQuadrature
    for e in range(ne):
        if fpm16 == 0:
            theta = ba * xg[e] + ab
            Z[e] = Emid + r * math.cos(theta) + 1j * r * aspect * math.sin(theta)
            jac = r * 1j * math.sin(theta) + r * aspect * math.cos(theta)
            W[e] = 0.25 * wg[e] * jac
        elif fpm16 == 2:
            zxe, zwe = zolotarev_point(ne, e + 1)
            Z[e] = zxe * r + Emid
            W[e] = zwe * r
        else:
            theta = math.pi - (math.pi / ne) / 2 - (math.pi / ne) * e
            Z[e] = Emid + r * math.cos(theta) + 1j * r * aspect * math.sin(theta)
            jac = r * 1j * math.sin(theta) + r * aspect * math.cos(theta)
            W[e] = (1.0 / (2 * ne)) * jac
    return Z, W


def feast_gcontour(Emid, r, fpm):
    """Full rotated ellipse (core/feast_tools.jl:286-371); Gauss = two half rules."""
    if fpm[7] == FEAST_UNINITIALIZED or fpm[7] <= 0:
        feastdefault(fpm)
    ne, fpm16, fpm18, fpm19 = fpm[7], fpm[15], fpm[17], fpm[18]
    Emid = complex(Emid)
    aspect = fpm18 * 0.01
    rot = (fpm19 / 180.0) * math.pi
    nr = r * (math.cos(rot) + 1j * math.sin(rot))
    ba, ab = -math.pi / 2, math.pi / 2
    Z = np.empty(ne, dtype=np.complex128)
    W = np.empty(ne, dtype=np.complex128)
    if fpm16 == 0:
        nu = ne // 2
        xu, wu = np.polynomial.legendre.leggauss(nu) if nu > 0 else (np.zeros(0), np.zeros(0))
        xl, wl = np.polynomial.legendre.leggauss(ne - nu)
        for e in range(nu):
            th = ba * xu[e] + ab
            Z[e] = Emid + nr * math.cos(th) + nr * 1j * aspect * math.sin(th)
            W[e] = 0.25 * wu[e] * (nr * 1j * math.sin(th) + nr * aspect * math.cos(th))
        for e in range(nu, ne):
            i = e - nu
            th = -ba * xl[i] - ab
            Z[e] = Emid + nr * math.cos(th) + nr * 1j * aspect * math.sin(th)
            W[e] = 0.25 * wl[i] * (nr * 1j * math.sin(th) + nr * aspect * math.cos(th))
    else:
        for e in range(ne):
            th = math.pi - (2 * math.pi / ne) / 2 - (2 * math.pi / ne) * e
            Z[e] = Emid + nr * math.cos(th) + nr * 1j * aspect * math.sin(th)
            W[e] = (1.0 / ne) * (nr * 1j * math.sin(th) + nr * aspect * math.cos(th))
    return Z, W


def feast_customcontour(Zne, fpm):
    """Trapezoid weights for user nodes (core/feast_tools.jl:378-398); sets fpm[2]."""
    Zne = np.asarray(Zne, dtype=np.complex128)
    ne = len(Zne)
    fpm[1] = ne
    W = np.empty(ne, dtype=np.complex128)
    for i in range(ne):
        W[i] = (Zne[(i + 1) % ne] - Zne[(i - 1) % ne]) / (2 * ne)
    return Zne, W


def feast_inside_contour(lam, Emin, Emax):
    """Closed interval (core/feast_tools.jl:619-621)."""
    return Emin <= lam <= Emax


def feast_inside_gcontour(lam, Emid, r, fpm=None):
    """Rotated-ellipse membership (core/feast_tools.jl:623-650)."""
    w = complex(lam) - complex(Emid)
    aspect, rot = 1.0, 0.0
    if fpm is not None and len(fpm) >= 19:
        if fpm[17] > 0:
            aspect = fpm[17] * 0.01
        if fpm[18] != 0:
            rot = (fpm[18] / 180.0) * math.pi
    if rot != 0.0:
        w *= complex(math.cos(-rot), math.sin(-rot))
    x = w.real / r
    y = w.imag / (r * aspect)
    return x * x + y * y <= 1.0


# --------------------------------------------------------------------------
# small helpers: core/feast_tools.jl:653-755, core/feast_aux.jl:84-257
# --------------------------------------------------------------------------
def feast_sort(lam, q, res, M):
    """Stable insertion sort ascending of the first M pairs (core/feast_tools.jl:653-681)."""
    idx = sorted(range(M), key=lambda i: lam[i])  # stable, same result as the insertion sort
    lam[:M] = np.asarray(lam)[idx]
    res[:M] = np.asarray(res)[idx]
    q[:, :M] = q[:, idx]


def feast_sort_general(lam, q, res, M):
    """Stable insertion sort by |lambda|^2 (core/feast_tools.jl:684-713)."""
    idx = sorted(range(M), key=lambda i: abs(lam[i]) ** 2)
    lam[:M] = np.asarray(lam)[idx]
    res[:M] = np.asarray(res)[idx]
    q[:, :M] = q[:, idx]


def feast_residual(A, B, lam, q, M):
    """res_j = ||A q_j - lam_j B q_j||_2 / max(|lam_j|,1) (core/feast_tools.jl:726-755)."""
    res = np.zeros(M)
    for j in range(M):
        r = A @ q[:, j] - lam[j] * (B @ q[:, j])
        res[j] = np.linalg.norm(r) / max(abs(lam[j]), 1.0)
    return res


def hermitian_part(src):
    """(S + S^H)/2 (core/feast_aux.jl:84-92)."""
    return 0.5 * (src + src.conj().T)


def qr_compress(src, ncols, rank_tol=None):
    """Pivoted-QR numerical range basis (core/feast_aux.jl:101-131).

    Returns (basis[:, :rank], rank).  Threshold = max(rank_tol, eps*max(dims))*|R11|.
    """
    real_dtype = np.zeros(1, dtype=src.dtype).real.dtype
    eps = np.finfo(real_dtype).eps
    if rank_tol is None:
        rank_tol = math.sqrt(eps)
    if ncols == 0:
        return src[:, :0].copy(), 0
    blk = np.array(src[:, :ncols])
    Qm, R, _ = sla.qr(blk, mode="economic", pivoting=True)
    rdiag = np.abs(np.diag(R))
    if rdiag.size == 0 or rdiag[0] == 0:
        return src[:, :0].copy(), 0
    thr = max(rank_tol, eps * max(blk.shape)) * rdiag[0]
    rank = 0
    for v in rdiag:
        if not v > thr:
            break
        rank += 1
    return Qm[:, :rank].copy(), rank


def reorder_by_interval(lam, vectors, Emin, Emax, M0):
    """Stable partition, inside first (core/feast_aux.jl:144-197). In place; returns M."""
    inside = [i for i in range(M0) if Emin <= lam[i] <= Emax]
    outside = [i for i in range(M0) if not (Emin <= lam[i] <= Emax)]
    perm = inside + outside
    lam[:M0] = np.asarray(lam)[perm]
    vectors[:, :M0] = vectors[:, perm]
    return len(inside)


def reorder_by_gcontour(lam, vectors, Emid, r, fpm, M0):
    """core/feast_aux.jl:208-257."""
    flags = [feast_inside_gcontour(lam[i], Emid, r, fpm) for i in range(M0)]
    perm = [i for i in range(M0) if flags[i]] + [i for i in range(M0) if not flags[i]]
    lam[:M0] = np.asarray(lam)[perm]
    vectors[:, :M0] = vectors[:, perm]
    return sum(flags)


def seeded_subspace(N, M0, seed=12345, complex_storage=True, dtype=np.float64):
    """Stand-in for _feast_seeded_subspace(_complex)! (core/feast_tools.jl:6-43).

    Julia's MersenneTwister(hash((N,M0))) stream is not reproducible outside Julia, so
    BOTH the oracle and the CUDA engine take Q0 from here: standard-normal REAL values,
    unit 2-norm columns (the reference also uses real values in complex storage).
    """
    rng = np.random.default_rng(seed)
    Q = rng.standard_normal((N, M0)).astype(dtype)
    nrm = np.linalg.norm(Q, axis=0)
    nrm[nrm == 0] = 1.0
    Q = Q / nrm
    return Q.astype(np.complex128 if dtype == np.float64 else np.complex64) if complex_storage else Q


@dataclass
class FeastResult:
    """FeastResult / FeastGeneralResult (core/feast_types.jl:85-108); arrays trimmed to M."""
    lambda_: np.ndarray
    q: np.ndarray
    M: int
    res: np.ndarray
    info: int
    epsout: float
    loop: int
    stats: dict = field(default_factory=dict)


# --------------------------------------------------------------------------
# shifted solves
# --------------------------------------------------------------------------
def _matmul(A, X):
    return A @ X


class _Shifted:
    """(z B - A) as a LinearOperator (sparse/feast_sparse.jl:9-27,133-148)."""

    def __init__(self, A, B, z):
        self.A, self.B, self.z = A, B, z
        n = A.shape[0]
        self.op = spla.LinearOperator((n, n), matvec=self.mv, dtype=np.complex128)

    def mv(self, x):
        bx = x if self.B is None else self.B @ x
        return self.z * bx - self.A @ x


def solve_shifted_iterative(A, B, z, rhs, tol, maxiter, restart, stats=None):
    """Per-column restarted GMRES + explicit gate (sparse/feast_sparse.jl:164-236).

    Krylov.jl gmres(restart=true, memory=max(restart,2), rtol=atol=tol, itmax=maxiter) is
    replaced by scipy gmres(rtol, atol, restart, maxiter=ceil(maxiter/restart)) with zero
    initial guess; then ||(zB-A)x-b|| <= 10*tol*max(||b||,1) is required.
    Returns (X, ok).
    """
    sh = _Shifted(A, B, z)
    X = np.zeros_like(rhs, dtype=np.complex128)
    mem = max(restart, 2)
    for j in range(rhs.shape[1]):
        b = rhs[:, j]
        cnt = [0]

        def cb(_):
            cnt[0] += 1
        x, info = spla.gmres(sh.op, b, rtol=tol, atol=tol, restart=mem,
                             maxiter=max(1, -(-maxiter // mem)), callback=cb, callback_type="pr_norm")
        if stats is not None:
            stats["krylov_iters"] = stats.get("krylov_iters", 0) + cnt[0]
        resn = np.linalg.norm(sh.mv(x) - b)
        if info != 0 or resn > 10 * tol * max(np.linalg.norm(b), 1.0):
            return X, False
        X[:, j] = x
    return X, True


def _factor(A, B, z, band_k=None):
    """lu(z*B - A) for dense (zgetrf), sparse (UMFPACK -> SuperLU) or -- band_k given -- banded inputs (the reference's banded route:
    shifted band in gbtrf layout, LAPACK.gbtrf!, banded/feast_banded.jl:216-237, 100-112, 678)."""
    if band_k is not None and sp.issparse(A):
        n, k = A.shape[0], int(band_k)
        Bm = sp.identity(n, dtype=np.complex128, format="csc") if B is None else B
        S = (z * Bm - A).tocsc().astype(np.complex128)
        ab = np.zeros((3 * k + 1, n), dtype=np.complex128)
        for d in range(-k, k + 1):
            if d >= 0:
                ab[2 * k - d, d:] = S.diagonal(d)
            else:
                ab[2 * k - d, :n + d] = S.diagonal(d)
        lu, piv, info = sla.lapack.zgbtrf(ab, k, k, overwrite_ab=True)
        if info != 0:
            raise np.linalg.LinAlgError(f"zgbtrf info = {info}")
        return ("band", (lu, piv, k))
    if sp.issparse(A):
        n = A.shape[0]
        Bm = sp.identity(n, dtype=np.complex128, format="csc") if B is None else B
        S = (z * Bm - A).tocsc().astype(np.complex128)
        return ("sparse", spla.splu(S))
    n = A.shape[0]
    S = (z * np.eye(n) - A) if B is None else (z * B - A)
    return ("dense", sla.lu_factor(S.astype(np.complex128)))


def _solve_factor(fac, rhs):
    kind, f = fac
    if kind == "band":                      # LAPACK.gbtrs!, banded/feast_banded.jl:141, 683
        lu, piv, k = f
        x, info = sla.lapack.zgbtrs(lu, k, k, np.asfortranarray(rhs, dtype=np.complex128), piv)
        return x
    if kind == "sparse":
        return f.solve(np.ascontiguousarray(rhs, dtype=np.complex128))
    return sla.lu_solve(f, rhs)


# --------------------------------------------------------------------------
# H-RR: dense/feast_dense.jl:78-351, sparse/feast_sparse.jl:246-499, banded/feast_banded.jl:561-823
# --------------------------------------------------------------------------
def feast_hrr(A, B, Emin, Emax, M0, fpm, Q0=None, solver="direct", solver_tol=0.0,
              solver_maxiter=500, solver_restart=30, filter="reference", contour=None,
              node_range=None, band_k=None, node_pool=None):
    """Hermitian FEAST with QR-compress + Rayleigh-Ritz.

    node_pool: callable rhs -> sum_e 2 w_e (z_e B - A)^-1 rhs evaluated by workers that each own one quadrature node and its cached
        factorisation -- the reference's :threads backend (`Threads.@threads for e in 1:ne`, parallel/feast_parallel.jl:586); direct
        solves with the plain half-contour sum only (filter="reference", or a real pencil with a real basis).

    band_k: half-bandwidth of a banded pencil given in sparse form -> the per-node factorisations are LAPACK band LUs (zgbtrf/zgbtrs),
        the reference's banded route (banded/feast_banded.jl:561-823); everything else is unchanged.

    A, B: dense ndarray or scipy sparse, Hermitian (B may be None = identity).
    filter="reference": accumulate the COMPLEX half-contour sum 2*w*Y exactly as the
        reference does (dense/feast_dense.jl:231, sparse/feast_sparse.jl:369).
    filter="true": additionally take the real part of the accumulated filter, i.e.
        rho(A) = Re g(A); exact for real-symmetric pencils with a real basis.
        For complex Hermitian pencils it adds the conjugate-node solves.
    Returns FeastResult with complex eigenvectors (FeastResult{T,Complex{T}}).
    """
    N = A.shape[0]
    feastdefault(fpm)
    check_feast_srci_input(N, M0, Emin, Emax, fpm)
    if solver == "iterative":
        solver = "gmres"
    if solver not in ("direct", "gmres"):
        raise ValueError("Unsupported solver")
    tol_value = 10.0 ** (-fpm[2]) if solver_tol == 0.0 else float(solver_tol)
    is_real_pencil = (not np.iscomplexobj(A)) and (B is None or not np.iscomplexobj(B))
    Ac = A.astype(np.complex128)
    Bc = None if B is None else B.astype(np.complex128)

    Q_basis = seeded_subspace(N, M0) if Q0 is None else np.array(Q0, dtype=np.complex128)
    # real-symmetric pencil + real basis + true filter: everything stays real-valued
    real_mode = filter == "true" and is_real_pencil and np.abs(Q_basis.imag).max() == 0.0
    Zne, Wne = contour if contour is not None else feast_contour(Emin, Emax, fpm)
    fac_cache = [None] * len(Zne)
    fac_cache_c = [None] * len(Zne)
    maxloop = fpm[3]
    eps_tol = feast_tolerance(fpm)
    epsout = math.inf
    info = SUCCESS
    loop_count = 0
    M_found = 0
    active = M0
    lam_vec = np.zeros(M0)
    res_vec = np.zeros(M0)
    solutions = np.zeros((N, M0), dtype=np.complex128)
    stats = {"krylov_iters": 0, "solves": 0}

    for loop_idx in range(maxloop + 1):
        loop_count = loop_idx
        Q_proj = np.zeros((N, M0), dtype=np.complex128)
        basis = Q_basis[:, :active]
        rhs = basis.copy() if Bc is None else Bc @ basis
        failed = False
        pooled = node_pool is not None and solver == "direct" and (filter == "reference" or real_mode)
        if pooled:
            Q_proj[:, :active] = node_pool(rhs)
            stats["solves"] += len(Zne)
        for e, z in enumerate(() if pooled else Zne):
            weight = 2 * Wne[e]
            if solver == "direct":
                if fac_cache[e] is None:
                    fac_cache[e] = _factor(Ac, Bc, z, band_k)
                Y = _solve_factor(fac_cache[e], rhs)
            else:
                Y, ok = solve_shifted_iterative(Ac, Bc, z, rhs, tol_value, solver_maxiter, solver_restart, stats)
                if not ok:
                    info = ERR_NO_CONV
                    failed = True
                    break
            stats["solves"] += 1
            if filter == "reference" or real_mode:
                Q_proj[:, :active] += weight * Y
            else:
                # true filter for a complex pencil/basis: w*Y(z) + conj(w)*Y(conj z)
                zc = np.conj(z)
                if solver == "direct":
                    if fac_cache_c[e] is None:
                        fac_cache_c[e] = _factor(Ac, Bc, zc, band_k)
                    Yc = _solve_factor(fac_cache_c[e], rhs)
                else:
                    Yc, ok = solve_shifted_iterative(Ac, Bc, zc, rhs, tol_value, solver_maxiter, solver_restart, stats)
                    if not ok:
                        info = ERR_NO_CONV
                        failed = True
                        break
                Q_proj[:, :active] += Wne[e] * Y + np.conj(Wne[e]) * Yc
        if failed:
            break
        if real_mode:
            # rho(A) = Re g(A) applied to a real basis of a real pencil: keep the real part
            Q_proj = Q_proj.real.astype(np.complex128)

        if real_mode:
            Qr, rank = qr_compress(np.ascontiguousarray(Q_proj.real), active)
            Qr = Qr.astype(np.complex128)
        else:
            Qr, rank = qr_compress(Q_proj, active)
        if rank == 0:
            info = ERR_NO_CONV
            break
        Sq = hermitian_part(Qr.conj().T @ (Ac @ Qr))
        if Bc is None:
            Aq = np.eye(rank, dtype=np.complex128)
        else:
            Aq = hermitian_part(Qr.conj().T @ (Bc @ Qr))
        try:
            if real_mode:
                lam_red, v_red = sla.eigh(Sq.real, Aq.real)
            else:
                lam_red, v_red = sla.eigh(Sq, Aq)
        except (np.linalg.LinAlgError, sla.LinAlgError):
            w, v_red = sla.eig(Sq, Aq)
            lam_red = w.real
        solutions[:, :rank] = Qr @ v_red
        lam_vec[:rank] = lam_red
        M = reorder_by_interval(lam_vec, solutions, Emin, Emax, rank)
        if M == 0:
            info = ERR_NO_CONV
            break
        for j in range(M):
            nrm = np.linalg.norm(solutions[:, j])
            if nrm > 0:
                solutions[:, j] /= nrm
        max_res = 0.0
        for j in range(M):
            x = solutions[:, j]
            rv = Ac @ x - lam_vec[j] * (x if Bc is None else Bc @ x)
            res_vec[j] = np.linalg.norm(rv) / max(abs(lam_vec[j]), 1.0)
            max_res = max(max_res, res_vec[j])
        epsout = max_res
        M_found = M
        if epsout <= eps_tol:
            break
        if loop_idx == maxloop:
            info = ERR_NO_CONV
            break
        active = rank
        Q_basis[:, :active] = solutions[:, :active]
    if M_found == 0:
        info = ERR_NO_CONV
    return FeastResult(lam_vec[:M_found].copy(), solutions[:, :M_found].copy(), M_found,
                       res_vec[:M_found].copy(), info, epsout, loop_count, stats)


def complex_to_real_result(r: FeastResult) -> FeastResult:
    """_complex_to_real_result: real.(q) (dense/feast_dense.jl:372-387)."""
    return FeastResult(r.lambda_.copy(), np.ascontiguousarray(r.q.real), r.M, r.res.copy(), r.info, r.epsout, r.loop, r.stats)


# reference-named wrappers -------------------------------------------------
def feast_syev(A, Emin, Emax, M0, fpm, **kw):
    """dense/feast_dense.jl:776-797"""
    return complex_to_real_result(feast_hrr(np.asarray(A, dtype=float), None, Emin, Emax, M0, fpm, **kw))


def feast_sygv(A, B, Emin, Emax, M0, fpm, **kw):
    """dense/feast_dense.jl:356-370"""
    return complex_to_real_result(feast_hrr(np.asarray(A, dtype=float), np.asarray(B, dtype=float), Emin, Emax, M0, fpm, **kw))


def feast_heev(A, Emin, Emax, M0, fpm, **kw):
    """dense/feast_dense.jl:390-400"""
    return feast_hrr(np.asarray(A, dtype=complex), None, Emin, Emax, M0, fpm, **kw)


def feast_hegv(A, B, Emin, Emax, M0, fpm, **kw):
    """dense/feast_dense.jl:799-810"""
    return feast_hrr(np.asarray(A, dtype=complex), np.asarray(B, dtype=complex), Emin, Emax, M0, fpm, **kw)


def feast_scsrev(A, Emin, Emax, M0, fpm, **kw):
    """sparse/feast_sparse.jl:1516-1529 (real symmetric sparse, standard)."""
    return complex_to_real_result(feast_hrr(sp.csc_matrix(A), None, Emin, Emax, M0, fpm, **kw))


def feast_scsrgv(A, B, Emin, Emax, M0, fpm, **kw):
    """sparse/feast_sparse.jl:713-731"""
    return complex_to_real_result(feast_hrr(sp.csc_matrix(A), sp.csc_matrix(B), Emin, Emax, M0, fpm, **kw))


def feast_hcsrev(A, Emin, Emax, M0, fpm, **kw):
    """sparse/feast_sparse.jl:759-788"""
    return feast_hrr(sp.csc_matrix(A, dtype=complex), None, Emin, Emax, M0, fpm, **kw)


def feast_hcsrgv(A, B, Emin, Emax, M0, fpm, **kw):
    """sparse/feast_sparse.jl:815-831"""
    return feast_hrr(sp.csc_matrix(A, dtype=complex), sp.csc_matrix(B, dtype=complex), Emin, Emax, M0, fpm, **kw)


# --------------------------------------------------------------------------
# banded storage helpers: banded/feast_banded.jl:205-314,423-483,1286-1318
# --------------------------------------------------------------------------
def full_to_banded(A, k):
    """Upper symmetric/Hermitian band (k+1) x n, diagonal in row k (0-based) (banded/feast_banded.jl:423-440)."""
    n = A.shape[0]
    AB = np.zeros((k + 1, n), dtype=A.dtype)
    for j in range(n):
        for i in range(max(0, j - k), j + 1):
            AB[k + i - j, j] = A[i, j]
    return AB


def banded_to_full(AB, k, hermitian=True):
    n = AB.shape[1]
    A = np.zeros((n, n), dtype=AB.dtype)
    for j in range(n):
        for i in range(max(0, j - k), j + 1):
            A[i, j] = AB[k + i - j, j]
            if i != j:
                A[j, i] = np.conj(AB[k + i - j, j]) if hermitian else AB[k + i - j, j]
    return A


def full_to_general_banded(A, k):
    """General band (2k+1) x n, diagonal in row k (banded/feast_banded.jl:1304-1318)."""
    n = A.shape[0]
    AB = np.zeros((2 * k + 1, n), dtype=A.dtype)
    for j in range(n):
        for i in range(max(0, j - k), min(n, j + k + 1)):
            AB[k + i - j, j] = A[i, j]
    return AB


def general_banded_to_full(AB, k):
    n = AB.shape[1]
    A = np.zeros((n, n), dtype=AB.dtype)
    for j in range(n):
        for i in range(max(0, j - k), min(n, j + k + 1)):
            A[i, j] = AB[k + i - j, j]
    return A


def symmetric_banded_matvec(AB, k, x):
    """banded/feast_banded.jl:239-259 (symmetric, NOT conjugated: real band storage)."""
    n = AB.shape[1]
    y = np.zeros(n, dtype=np.result_type(AB.dtype, x.dtype))
    for j in range(n):
        for i in range(max(0, j - k), j + 1):
            v = AB[k + i - j, j]
            y[i] += v * x[j]
            if i != j:
                y[j] += v * x[i]
    return y


# --------------------------------------------------------------------------
# S-MOM: kernel/feast_kernel.jl:7-293 driven as in banded/feast_banded.jl:9-186
# --------------------------------------------------------------------------
def feast_smom(A, B, Emin, Emax, M0, fpm, Q0=None, contour=None):
    """Real-symmetric moment FEAST (feast_srci! + a direct-solve driver).

    A, B real symmetric (dense or sparse; B None = identity).  Returns real eigenvectors,
    sorted ascending at exit (feast_sort!, kernel/feast_kernel.jl:260).  The residual
    ignores B exactly like the reference (kernel/feast_kernel.jl:250) unless B is None.
    """
    N = A.shape[0]
    feastdefault(fpm)
    if N <= 0:
        return FeastResult(np.zeros(0), np.zeros((N, 0)), 0, np.zeros(0), ERR_N, 0.0, 0)
    if M0 <= 0 or M0 > N:
        return FeastResult(np.zeros(0), np.zeros((N, 0)), 0, np.zeros(0), ERR_M0, 0.0, 0)
    if Emin >= Emax:
        return FeastResult(np.zeros(0), np.zeros((N, 0)), 0, np.zeros(0), ERR_EMIN_EMAX, 0.0, 0)
    Zne, Wne = contour if contour is not None else feast_contour(Emin, Emax, fpm)
    ne = len(Zne)
    Q0 = seeded_subspace(N, M0, complex_storage=False) if Q0 is None else np.array(np.real(Q0), dtype=float)
    Ac = A.astype(np.complex128)
    Bc = None if B is None else B.astype(np.complex128)
    fac = [None] * ne
    loop = 0
    lam = np.zeros(M0)
    q = np.zeros((N, M0))
    res = np.zeros(M0)
    eps_tol = feast_tolerance(fpm)
    maxloop = fpm[3]
    info = SUCCESS
    epsout = 0.0
    M = 0
    while True:
        Q_proj = np.zeros((N, M0), dtype=np.complex128)
        zAq = np.zeros((M0, M0), dtype=np.complex128)
        zSq = np.zeros((M0, M0), dtype=np.complex128)
        rhs = Q0.astype(np.complex128) if Bc is None else Bc @ Q0
        for e in range(ne):
            if fac[e] is None:
                fac[e] = _factor(Ac, Bc, Zne[e])
            Y = _solve_factor(fac[e], rhs)
            w = 2 * Wne[e]
            Q_proj += w * Y
            mom = Q0.T @ Y
            zAq += w * mom
            zSq += Zne[e] * w * mom
        Aq, Sq = zAq.real.copy(), zSq.real.copy()
        try:
            wv, V = sla.eig(Sq, Aq)
        except Exception:
            info = ERR_LAPACK
            break
        lam[:] = wv.real
        V = V.real
        q[:, :] = Q_proj.real @ V
        M = reorder_by_interval(lam, q, Emin, Emax, M0)
        if M == 0:
            info = ERR_NO_CONV
            break
        for j in range(M):
            rv = (A @ q[:, j]) - lam[j] * q[:, j]
            res[j] = np.linalg.norm(rv) / max(abs(lam[j]), 1.0)
        epsout = float(res[:M].max())
        if epsout <= eps_tol or loop >= maxloop:
            feast_sort(lam, q, res, M)
            break
        loop += 1
        Q0 = q[:, :M0].copy()
    return FeastResult(lam[:M].copy(), q[:, :M].copy(), M, res[:M].copy(), info, epsout, loop)


def feast_hmom(A, B, Emin, Emax, M0, fpm, Q0=None, contour=None, max_sweeps=None):
    """Complex-Hermitian moment FEAST: feast_hrci! (kernel/feast_kernel.jl:397-644) with a direct-solve driver.

    Restated AS WRITTEN: the half-contour sums use 2 w_e Y without a Hermitian part (:516-524), so the moments are
    Q0^H g(A) Q0 and Q0^H h(A) Q0 with the complex g(x) = sum 2w/(z-x), h(x) = sum 2wz/(z-x) = c + x g(x), c = sum 2w; the Ritz
    values real(eig(zSq, zAq)) equal x + Re(c / g(x)) -- exact only at the centre of the interval.  The reference has no internal
    caller and tests only the INIT handshake of this routine.  Returns (FeastResult, zAq, zSq of the last sweep)."""
    N = A.shape[0]
    feastdefault(fpm)
    Zne, Wne = contour if contour is not None else feast_contour(Emin, Emax, fpm)
    ne = len(Zne)
    Q0 = seeded_subspace(N, M0) if Q0 is None else np.array(Q0, dtype=np.complex128)
    Ac = A.astype(np.complex128)
    Bc = None if B is None else B.astype(np.complex128)
    fac = [None] * ne
    lam, q, res = np.zeros(M0), np.zeros((N, M0), dtype=np.complex128), np.zeros(M0)
    eps_tol, maxloop = feast_tolerance(fpm), fpm[3]
    info, epsout, loop, M, sweeps = SUCCESS, 0.0, 0, 0, 0
    zAq = zSq = None
    while True:
        Q_proj = np.zeros((N, M0), dtype=np.complex128)
        zAq = np.zeros((M0, M0), dtype=np.complex128)
        zSq = np.zeros((M0, M0), dtype=np.complex128)
        rhs = Q0 if Bc is None else Bc @ Q0
        for e in range(ne):
            if fac[e] is None:
                fac[e] = _factor(Ac, Bc, Zne[e])
            Y = _solve_factor(fac[e], rhs)
            w = 2 * Wne[e]
            Q_proj += w * Y
            mom = Q0.conj().T @ Y
            zAq += w * mom
            zSq += w * Zne[e] * mom
        sweeps += 1
        wv, V = sla.eig(zSq, zAq)
        lam[:] = wv.real
        q[:, :] = Q_proj @ V
        M = reorder_by_interval(lam, q, Emin, Emax, M0)
        if M == 0:
            info = ERR_NO_CONV
            break
        for j in range(M):
            rv = (Ac @ q[:, j]) - lam[j] * q[:, j]
            res[j] = np.linalg.norm(rv) / max(abs(lam[j]), 1.0)
        epsout = float(res[:M].max())
        if epsout <= eps_tol or loop >= maxloop or (max_sweeps is not None and sweeps >= max_sweeps):
            if max_sweeps is None or sweeps < max_sweeps:
                feast_sort(lam, q, res, M)
            break
        loop += 1
        Q0 = q[:, :M0].copy()
    return FeastResult(lam[:M].copy(), q[:, :M].copy(), M, res[:M].copy(), info, epsout, loop), zAq, zSq


def feast_sbgv(AB, BB, kla, klb, Emin, Emax, M0, fpm, **kw):
    """banded/feast_banded.jl:9-186: real symmetric banded generalized via S-MOM."""
    A = banded_to_full(np.asarray(AB, dtype=float), kla, hermitian=False)
    B = banded_to_full(np.asarray(BB, dtype=float), klb, hermitian=False)
    return feast_smom(sp.csc_matrix(A), sp.csc_matrix(B), Emin, Emax, M0, fpm, **kw)


def feast_sbev(AB, kla, Emin, Emax, M0, fpm, **kw):
    """banded/feast_banded.jl:1410-1420: B = identity band."""
    A = banded_to_full(np.asarray(AB, dtype=float), kla, hermitian=False)
    return feast_smom(sp.csc_matrix(A), sp.identity(A.shape[0], format="csc"), Emin, Emax, M0, fpm, **kw)


def feast_hbev(AB, kla, Emin, Emax, M0, fpm, **kw):
    """banded/feast_banded.jl:326-383 -> H-RR with band storage (banded/feast_banded.jl:561-823)."""
    A = banded_to_full(np.asarray(AB, dtype=complex), kla, hermitian=True)
    return feast_hrr(sp.csc_matrix(A), None, Emin, Emax, M0, fpm, **kw)


def feast_hbgv(AB, BB, kla, klb, Emin, Emax, M0, fpm, **kw):
    """banded/feast_banded.jl:385-421"""
    A = banded_to_full(np.asarray(AB, dtype=complex), kla, hermitian=True)
    B = banded_to_full(np.asarray(BB, dtype=complex), klb, hermitian=True)
    return feast_hrr(sp.csc_matrix(A), sp.csc_matrix(B), Emin, Emax, M0, fpm, **kw)


# --------------------------------------------------------------------------
# G-RCI: kernel/feast_kernel.jl:646-962 + drivers dense:402-593, sparse:873-1006
# --------------------------------------------------------------------------
def feast_general(A, B, Emid, r, M0, fpm, Q0=None, solver="direct", solver_tol=0.0,
                  solver_maxiter=500, solver_restart=30, residual="reference", contour=None):
    """General (non-Hermitian) FEAST: full contour, one-sided Rayleigh-Ritz, no QR.

    residual="reference": ||A x - lam x|| / max(|lam|,1) even when B != I
        (kernel/feast_kernel.jl:900-906 -- B is ignored, so generalized problems only stop
        at fpm[4]); residual="true": ||A x - lam B x||.
    Exit sorts the M inside pairs by |lam| (kernel/feast_kernel.jl:918).
    """
    N = A.shape[0]
    feastdefault(fpm)
    check_feast_grci_input(N, M0, Emid, r, fpm)
    Emid = complex(Emid)
    if solver == "iterative":
        solver = "gmres"
    tol_value = 10.0 ** (-fpm[2]) if solver_tol == 0.0 else float(solver_tol)
    Ac = A.astype(np.complex128)
    Bc = None if B is None else B.astype(np.complex128)
    Zne, Wne = contour if contour is not None else feast_gcontour(Emid, r, fpm)
    ne = len(Zne)
    Qb = seeded_subspace(N, M0) if Q0 is None else np.array(Q0, dtype=np.complex128)
    fac = {}
    loop = 0
    info = SUCCESS
    epsout = 0.0
    lam = np.zeros(M0, dtype=np.complex128)
    q = np.zeros((N, M0), dtype=np.complex128)
    res = np.zeros(M0)
    M = 0
    eps_tol = feast_tolerance(fpm)
    maxloop = fpm[3]
    stats = {"krylov_iters": 0, "solves": 0}
    while True:
        q[:] = 0
        rhs = Qb.copy() if Bc is None else Bc @ Qb
        failed = False
        for e in range(ne):
            z = Zne[e]
            if solver == "direct":
                if z not in fac:
                    fac[z] = _factor(Ac, Bc, z)
                Y = _solve_factor(fac[z], rhs)
            else:
                Y, ok = solve_shifted_iterative(Ac, Bc, z, rhs, tol_value, solver_maxiter, solver_restart, stats)
                if not ok:
                    info = ERR_NO_CONV
                    failed = True
                    break
            stats["solves"] += 1
            q += Wne[e] * Y
        if failed:
            break
        Sq = q.conj().T @ (q if Bc is None else Bc @ q)
        Aq = q.conj().T @ (Ac @ q)
        try:
            lam_red, v_red = sla.eig(Aq, Sq)
        except Exception:
            info = ERR_LAPACK
            break
        # Julia's eigen(A,B) sorts lexicographically by (real, imag) (LinearAlgebra eigsortby)
        order = sorted(range(M0), key=lambda i: (lam_red[i].real, lam_red[i].imag))
        lam_red = lam_red[order]
        v_red = v_red[:, order]
        X = q @ v_red
        lam[:] = lam_red
        M = reorder_by_gcontour(lam, X, Emid, r, fpm, M0)
        if M == 0:
            info = ERR_NO_CONV
            break
        nrm = np.linalg.norm(X, axis=0)
        nrm[nrm == 0] = 1.0
        X = X / nrm
        q[:, :] = X
        for j in range(M):
            bx = q[:, j] if (Bc is None or residual == "reference") else Bc @ q[:, j]
            rv = Ac @ q[:, j] - lam[j] * bx
            res[j] = np.linalg.norm(rv) / max(abs(lam[j]), 1.0)
        epsout = float(res[:M].max())
        if epsout <= eps_tol or loop >= maxloop:
            feast_sort_general(lam, q, res, M)
            break
        loop += 1
        Qb = q.copy()
    return FeastResult(lam[:M].copy(), q[:, :M].copy(), M, res[:M].copy(), info, epsout, loop, stats)


# --------------------------------------------------------------------------
# node sharding rule: parallel/feast_mpi.jl:36-43, parallel/feast_parallel.jl:433-447
# --------------------------------------------------------------------------
def node_partition(ne, nranks, rank):
    """Contiguous block distribution of quadrature nodes; returns (start, count), 0-based."""
    base, rem = divmod(ne, nranks)
    start = rank * base + min(rank, rem)
    count = base + (1 if rank < rem else 0)
    return start, count


def partial_accumulate(A, B, Zne, Wne, Q0, start, count, factor=2.0):
    """One rank's share of the filter: sum_{e in block} factor*w_e (z_e B - A)^{-1} B Q0
    (parallel/feast_mpi.jl:796-909 structure); the ranks' results are Allreduce-summed."""
    Ac = A.astype(np.complex128)
    Bc = None if B is None else B.astype(np.complex128)
    Q0 = np.asarray(Q0, dtype=np.complex128)
    rhs = Q0 if Bc is None else Bc @ Q0
    acc = np.zeros_like(Q0)
    for e in range(start, start + count):
        acc += factor * Wne[e] * _solve_factor(_factor(Ac, Bc, Zne[e]), rhs)
    return acc


# --------------------------------------------------------------------------
# synthetic workloads of BASELINE.json / SURVEY §8d (shared by tests and bench)
# --------------------------------------------------------------------------
def laplacian_1d(n):
    return sp.diags([-np.ones(n - 1), 2 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1], format="csr")


def laplacian_3d(nx, ny=None, nz=None):
    """7-point Dirichlet Laplacian, Kronecker sum of tridiag(-1,2,-1)."""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    Tx, Ty, Tz = laplacian_1d(nx), laplacian_1d(ny), laplacian_1d(nz)
    Ix, Iy, Iz = sp.identity(nx), sp.identity(ny), sp.identity(nz)
    A = sp.kron(sp.kron(Tx, Iy), Iz) + sp.kron(sp.kron(Ix, Ty), Iz) + sp.kron(sp.kron(Ix, Iy), Tz)
    A = A.tocsr()
    A.sort_indices()
    return A


def laplacian_3d_eigs(nx, ny=None, nz=None):
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    mu = lambda n: 2 - 2 * np.cos(np.arange(1, n + 1) * np.pi / (n + 1))
    e = (mu(nx)[:, None, None] + mu(ny)[None, :, None] + mu(nz)[None, None, :]).ravel()
    return np.sort(e)


def subspace_angle(X, Y):
    """Largest principal angle (radians) between span(X) and span(Y)."""
    Qx, _ = np.linalg.qr(X)
    Qy, _ = np.linalg.qr(Y)
    if Qx.shape[1] != Qy.shape[1]:
        return math.pi / 2
    # sin(theta_max) = || (I - Qx Qx^H) Qy ||_2
    R = Qy - Qx @ (Qx.conj().T @ Qy)
    s = np.linalg.norm(R, 2)
    return float(math.asin(min(1.0, s)))
