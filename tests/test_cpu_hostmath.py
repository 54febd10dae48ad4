"""CPU tests of the engine's host-side dense routines (csrc/host_math.hpp: LU solve, Hessenberg/shifted-QR eigen-solver, the
rank-revealing reduced-pencil solver behind feastcuda_eig_general) against LAPACK, through a g++-built harness -- no GPU."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest
import scipy.linalg as sla

import feast_oracle as fo

ROOT = Path(__file__).resolve().parents[1]
SRC = ROOT / "tests" / "helpers" / "hostmath_harness.cpp"
OUT = ROOT / "tests" / "helpers" / "libhostmath_harness.so"
_vp = C.c_void_p


@pytest.fixture(scope="module")
def hm():
    deps = [SRC, ROOT / "feastkit.jl_b200" / "csrc" / "host_math.hpp"]
    if not OUT.exists() or any(d.stat().st_mtime > OUT.stat().st_mtime for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", str(OUT), str(SRC)], check=True)
    lib = C.CDLL(str(OUT))
    for f in (lib.hm_pencil_eig, lib.hm_complex_eig, lib.hm_lu_solve):
        f.restype = C.c_int
    return lib


def _f(a):
    return np.asfortranarray(a, dtype=np.complex128)


def _ptr(a):
    return a.ctypes.data_as(_vp)


def pencil_eig(hm, S, B):
    n = S.shape[0]
    S, B = _f(S), _f(B)
    lam, V = np.zeros(n, dtype=np.complex128), np.zeros((n, n), dtype=np.complex128, order="F")
    rank = hm.hm_pencil_eig(C.c_long(n), _ptr(S), _ptr(B), _ptr(lam), _ptr(V))
    return rank, lam, V


def _match(got, want, tol):
    want = list(want)
    assert len(got) == len(want)
    for g in got:
        j = int(np.argmin([abs(g - w) for w in want]))
        assert abs(g - want[j]) <= tol * max(1.0, abs(want[j])), (g, want[j])
        want.pop(j)


@pytest.mark.parametrize("n", [1, 2, 3, 8, 33, 64])
def test_complex_eig_matches_lapack(hm, n):
    rng = np.random.default_rng(n)
    A = _f(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    lam, V = np.zeros(n, dtype=np.complex128), np.zeros((n, n), dtype=np.complex128, order="F")
    assert hm.hm_complex_eig(C.c_long(n), _ptr(A), _ptr(lam), _ptr(V)) == 0
    _match(lam, np.linalg.eigvals(A), 1e-10)
    assert np.linalg.norm(A @ V - V * lam) <= 1e-10 * np.linalg.norm(A) * n
    assert np.allclose(np.linalg.norm(V, axis=0), 1.0)
    # Hermitian and real non-symmetric inputs (conjugate pairs), and a matrix with a repeated eigenvalue
    for M in (A + A.conj().T, rng.standard_normal((n, n)).astype(complex), np.diag(np.repeat([1.0, 2.0], [n - n // 2, n // 2])).astype(complex)):
        M = _f(M)
        assert hm.hm_complex_eig(C.c_long(n), _ptr(M), _ptr(lam), _ptr(V)) == 0
        _match(lam, np.linalg.eigvals(M), 1e-9)


def test_lu_solve_matches_lapack_and_flags_singular(hm):
    rng = np.random.default_rng(0)
    n = 17
    M = _f(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    Bm = _f(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    X = np.zeros((n, n), dtype=np.complex128, order="F")
    assert hm.hm_lu_solve(C.c_long(n), _ptr(M), _ptr(Bm), _ptr(X)) == 0
    assert np.linalg.norm(M @ X - Bm) <= 1e-12 * np.linalg.norm(M) * np.linalg.norm(X)
    Z = _f(np.zeros((n, n)))
    assert hm.hm_lu_solve(C.c_long(n), _ptr(Z), _ptr(Bm), _ptr(X)) == -1


@pytest.mark.parametrize("n", [2, 5, 12, 40])
def test_pencil_eig_regular_pencils_match_qz(hm, n):
    rng = np.random.default_rng(100 + n)
    S = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    B = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    rank, lam, V = pencil_eig(hm, S, B)
    assert rank == n
    _match(lam, sla.eigvals(S, B), 1e-8)
    assert np.linalg.norm(S @ V - B @ V * lam) <= 1e-9 * (np.linalg.norm(S) + np.linalg.norm(B) * np.abs(lam).max())


def test_pencil_eig_deflates_the_null_space_of_feast_moment_pencils(hm):
    """Aq = Q0^T rho(A) Q0, Sq = Q0^T A rho(A) Q0 with k eigenvalues inside and M0 > k columns: rank k, the k Ritz values exact,
    the other M0 - k reported as +inf (LAPACK's QZ returns arbitrary values there)."""
    rng = np.random.default_rng(7)
    n, m0, k = 50, 14, 5
    U, _ = np.linalg.qr(rng.standard_normal((n, n)))
    ev = np.linspace(0.1, 4.0, n)
    A = (U * ev) @ U.T
    rho = np.where(np.arange(n) < k, 1.0, 1e-18)
    P = (U * rho) @ U.T
    Q0 = rng.standard_normal((n, m0))
    Aq, Sq = Q0.T @ P @ Q0, Q0.T @ A @ P @ Q0
    rank, lam, V = pencil_eig(hm, Sq, Aq)
    assert rank == k
    fin = np.isfinite(lam.real)
    assert fin.sum() == k and np.allclose(np.sort(lam[fin].real), ev[:k], atol=1e-9) and np.abs(lam[fin].imag).max() < 1e-9
    X = P @ Q0 @ V[:, fin]
    assert (np.linalg.norm(A @ X - X * lam[fin].real, axis=0) / np.linalg.norm(X, axis=0)).max() < 1e-9
    # the other M0 - k vectors are the coordinate vectors of the pivoted-out columns: with the k Ritz vectors they keep V regular,
    # so that q = Q_proj V still spans the whole filtered block in the next refinement loop
    assert np.allclose(np.sort(np.abs(V[:, ~fin]), axis=0)[-1], 1.0) and np.count_nonzero(V[:, ~fin]) == m0 - k
    assert np.linalg.matrix_rank(V) == m0


def test_pencil_eig_infinite_eigenvalues_and_bad_input(hm):
    rng = np.random.default_rng(3)
    S = rng.standard_normal((6, 6))
    B = rng.standard_normal((6, 6))
    B[:, 4:] = B[:, :2] @ rng.standard_normal((2, 2))                             # rank 4: two infinite eigenvalues, regular pencil
    rank, lam, V = pencil_eig(hm, S, B)
    w = sla.eigvals(S, B)
    assert rank == 4
    _match(lam[np.isfinite(lam)], w[np.isfinite(w)], 1e-8)
    for c in np.where(np.isfinite(lam))[0]:
        assert np.linalg.norm(S @ V[:, c] - lam[c] * (B @ V[:, c])) < 1e-9
    assert pencil_eig(hm, S, np.zeros((6, 6)))[0] == 0                            # B = 0: every eigenvalue infinite
    S[0, 0] = np.nan
    assert pencil_eig(hm, S, B)[0] == -1                                          # non-finite input is refused


def _orthonormalize(hm, Z, rank_tol=0.0):
    """The engine's orthonormalize() loop (csrc/feastcuda.cu) with NumPy standing in for the device Gram / row-transform
    kernels: Gram -> ortho_pass (host steering) -> Z <- Z T, until every kept column is accepted and the Gram is the identity."""
    n, ncols = Z.shape
    cur, done, thr_abs, eps = ncols, 0, -1.0, np.finfo(float).eps
    passes = 0
    for p in range(16):
        G = _f(Z.conj().T @ Z)
        passes += 1
        if p == 0:
            dmax = G.diagonal().real.max()
            if not dmax > 0:
                return Z[:, :0], 0, passes
            thr_abs = max(rank_tol, eps * max(n, ncols)) * np.sqrt(dmax)
        if done == cur and np.abs(G - np.eye(cur)).max() <= 1e-14:
            break
        T = np.zeros((cur, cur), dtype=np.complex128, order="F")
        out3 = (C.c_long * 3)()
        hm.hm_ortho_pass(C.c_long(cur), C.c_long(done), _ptr(G), C.c_double(thr_abs), C.c_double(1e-8), _ptr(T), out3)
        accepted, kept = int(out3[0]), int(out3[1])
        if accepted + kept == 0:
            return Z[:, :0], 0, passes
        Z = Z @ T[:, :accepted + kept]
        done, cur = accepted, accepted + kept
    return Z[:, :done], done, passes


def test_gram_driven_orthonormalisation_matches_pivoted_qr(hm):
    """K7 of the engine (_feast_qr_compress!, core/feast_aux.jl:101-131): same numerical rank and the same range as the
    pivoted-QR restatement in the oracle, orthonormal to 1e-14 -- on a well-conditioned block, on graded columns (condition 1e10)
    and on rank-deficient blocks including the reference's own fixture (test_allocation_helpers.jl:274-292)."""
    rng = np.random.default_rng(11)
    n = 300
    blocks = {
        "well": rng.standard_normal((n, 24)) + 1j * rng.standard_normal((n, 24)),
        "graded": (rng.standard_normal((n, 20)) + 1j * rng.standard_normal((n, 20))) * np.logspace(0, -10, 20),
        "deficient": (rng.standard_normal((n, 7)) + 1j * rng.standard_normal((n, 7))) @ (rng.standard_normal((7, 18)) + 0j),
        "filtered": np.linalg.qr(rng.standard_normal((n, n)))[0][:, :30] * np.r_[np.ones(9), np.full(21, 1e-17)] @ rng.standard_normal((30, 30)),
    }
    src = np.array([[1.0, 2.0, 0.0, 1.0e-15], [1.0j, 2.0j, 1.0, 1.0e-15j], [0, 0, 1.0j, 0], [0, 0, 0, 0]], dtype=complex)
    blocks["ka12"] = src
    for name, Zb in blocks.items():
        Zb = np.asarray(Zb, dtype=complex)
        rank_tol = np.sqrt(np.finfo(float).eps)                                   # the reference's default threshold
        Q, rank, passes = _orthonormalize(hm, Zb.copy(), rank_tol)
        Qo, rank_o = fo.qr_compress(Zb, Zb.shape[1])
        assert rank == rank_o, (name, rank, rank_o)
        assert np.abs(Q.conj().T @ Q - np.eye(rank)).max() < 1e-13, name
        assert fo.subspace_angle(Q, Qo) < 1e-7, name
        assert passes <= 6, (name, passes)
    assert _orthonormalize(hm, np.zeros((10, 3), dtype=complex))[1] == 0
