"""GPU parity at reduced size for the BASELINE.json configs that are not the bench workload (configs[0], [1], [3], [4]);
synthetic matrices as SURVEY.md §8d defines them, checked against analytic spectra and the oracle."""
import numpy as np
import pytest
import scipy.linalg as sla
import scipy.sparse as sp

import feast_oracle as fo

pytestmark = pytest.mark.gpu


def test_config0_readme_tridiagonal_n100():
    """configs[0]: README tridiagonal 1-D Laplacian n=100, feast(A,(0.5,1.5),M0=10), dense Float64.  19 eigenvalues lie in the
    interval (> M0 = 10), so the reference cannot converge: API smoke test; M0 = 30 is the parity case."""
    import feastcuda as fc
    n = 100
    A = fo.laplacian_1d(n).toarray()
    lam = 2 - 2 * np.cos(np.arange(1, n + 1) * np.pi / (n + 1))
    inside = lam[(lam >= 0.5) & (lam <= 1.5)]
    assert len(inside) == 19
    r10 = fc.feast(A, (0.5, 1.5), M0=10, fpm=fc.feastinit(), Q0=fo.seeded_subspace(n, 10, complex_storage=False))
    ro10 = fo.feast_syev(A, 0.5, 1.5, 10, fo.feastinit(), Q0=fo.seeded_subspace(n, 10), filter="true")
    assert r10.info == ro10.info == 5          # same failure mode as the reference's algorithm
    Q0 = fo.seeded_subspace(n, 30, complex_storage=False)
    r = fc.feast(A, (0.5, 1.5), M0=30, fpm=fc.feastinit(), Q0=Q0)
    ro = fo.feast_syev(A, 0.5, 1.5, 30, fo.feastinit(), Q0=Q0.astype(complex), filter="true")
    assert r.info == ro.info == 0 and r.M == ro.M == 19
    assert np.abs(np.sort(r.lambda_) - inside).max() < 1e-10 and r.res.max() < 1e-12
    assert fo.subspace_angle(r.q.astype(complex), ro.q.astype(complex)) < 1e-8
    rs = fc.feast(sp.csc_matrix(A), (0.5, 1.5), M0=30, fpm=fc.feastinit(), Q0=Q0)   # sparse route: multi-shift Lanczos
    assert rs.info == 0 and rs.M == 19 and np.abs(np.sort(rs.lambda_) - inside).max() < 1e-10


def test_config1_dense_householder_similar_reduced():
    """configs[1] at n=1024 (full size 8192): A = H D H, H = I - 2vv^T (seed 42), D = diag(linspace(0,100,n)), M0=128, 8 nodes."""
    import feastcuda as fc
    n, M0 = 1024, 128
    rng = np.random.default_rng(42)
    v = rng.standard_normal(n)
    v /= np.linalg.norm(v)
    d = np.linspace(0.0, 100.0, n)
    Dv = d * v
    A = np.diag(d) - 2 * np.outer(v, Dv) - 2 * np.outer(Dv, v) + 4 * (v @ Dv) * np.outer(v, v)
    A = 0.5 * (A + A.T)
    Emin, Emax = 50.0, 50.0 + 80.5 * (100.0 / (n - 1))
    inside = d[(d >= Emin) & (d <= Emax)]
    assert 70 <= len(inside) <= 90
    Q0 = fo.seeded_subspace(n, M0, complex_storage=False)
    r = fc.dfeast_syev(A, Emin, Emax, M0, fc.feastinit(), Q0=Q0)
    assert r.info == 0 and r.M == len(inside)
    assert np.abs(np.sort(r.lambda_) - inside).max() < 1e-10 * 100.0 and r.res.max() < 1e-12
    order = np.argsort(r.lambda_)
    idx = np.where((d >= Emin) & (d <= Emax))[0]
    Hcols = np.eye(n)[:, idx] - 2 * np.outer(v, v[idx])            # exact eigenvectors H e_i
    overlap = np.abs(np.sum(Hcols * r.q[:, order], axis=0))
    assert np.abs(overlap - 1.0).max() < 1e-8


def _fem_pair(nx, ny, nz, seed=7):
    def k1(n, h):
        return sp.diags([-np.ones(n - 1), 2 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1]) / h

    def m1(n, h):
        return h * sp.diags([np.ones(n - 1), 4 * np.ones(n), np.ones(n - 1)], [-1, 0, 1]) / 6
    hs = [1.0 / (n + 1) for n in (nx, ny, nz)]
    Ks = [k1(n, h) for n, h in zip((nx, ny, nz), hs)]
    Ms = [m1(n, h) for n, h in zip((nx, ny, nz), hs)]
    K = sp.kron(sp.kron(Ks[0], Ms[1]), Ms[2]) + sp.kron(sp.kron(Ms[0], Ks[1]), Ms[2]) + sp.kron(sp.kron(Ms[0], Ms[1]), Ks[2])
    Mass = sp.kron(sp.kron(Ms[0], Ms[1]), Ms[2])
    n = nx * ny * nz
    phi = np.random.default_rng(seed).uniform(0, 2 * np.pi, n)
    D = sp.diags(np.exp(1j * phi))
    A = (D @ K @ D.conj()).tocsc()
    B = (D @ Mass @ D.conj()).tocsc()
    A = ((A + A.conj().T) * 0.5).tocsc()
    B = ((B + B.conj().T) * 0.5).tocsc()
    lam = []
    for n_, h in zip((nx, ny, nz), hs):
        th = np.arange(1, n_ + 1) * np.pi / (n_ + 1)
        lam.append(((2 - 2 * np.cos(th)) / h) / (h * (4 + 2 * np.cos(th)) / 6))
    w = np.sort((lam[0][:, None, None] + lam[1][None, :, None] + lam[2][None, None, :]).ravel())
    return A, B, w


def test_config3_hermitian_generalized_fem_pair_reduced():
    """configs[3] at 8x8x6 (full size 100x100x50): complex Hermitian stiffness/mass pair under a diagonal gauge, zfeast_hcsrgv!."""
    import feastcuda as fc
    A, B, w = _fem_pair(8, 8, 6)
    n = A.shape[0]
    M0 = 24
    want = 7
    assert w[want] - w[want - 1] > 1e-6 * w[want]
    Emin, Emax = 0.0, 0.5 * (w[want - 1] + w[want])
    Q0 = fo.seeded_subspace(n, M0)
    fpm = fc.feastinit()
    fpm[3] = 60
    r = fc.zfeast_hcsrgv(A, B, Emin, Emax, M0, list(fpm), Q0=Q0, solver_tol=1e-12, solver_maxiter=4000, ritz_guess=True, inner_rel=1e-9)
    ro = fo.feast_hcsrgv(A, B, Emin, Emax, M0, list(fpm), Q0=Q0)
    assert r.info == ro.info == 0 and r.M == ro.M == want
    assert np.abs(np.sort(r.lambda_) - w[:want]).max() < 1e-10 * w[want]
    assert r.res.max() < 1e-12
    assert fo.subspace_angle(r.q, ro.q) < 1e-8


def test_config4_general_complex_toeplitz_kronecker_reduced():
    """configs[4] at 5x5x8 (full size 50x50x100): Kronecker sums of non-symmetric complex Toeplitz tridiagonals,
    B = I + eps * (Kronecker sum with the same b/c ratio); analytic eigenvalues; circular contour, pzifeast_gcsrgv!."""
    import feastcuda as fc
    dims = (5, 5, 8)
    coef = [(0.4 + 0.1j, 1.0 + 0.05j, 0.9 - 0.05j), (0.3 - 0.1j, 0.8 + 0.1j, 0.75 + 0.05j), (0.5 + 0.2j, 0.6 - 0.05j, 0.65 + 0.02j)]
    eps = 0.05

    def toe(n, a, b, c):
        return sp.diags([b * np.ones(n - 1), a * np.ones(n), c * np.ones(n - 1)], [-1, 0, 1])
    I = [sp.identity(n) for n in dims]
    T = [toe(n, *abc) for n, abc in zip(dims, coef)]
    S = [toe(n, 0.0, abc[1], abc[2]) for n, abc in zip(dims, coef)]      # same b/c ratio -> same eigenvectors
    ksum = lambda X: sp.kron(sp.kron(X[0], I[1]), I[2]) + sp.kron(sp.kron(I[0], X[1]), I[2]) + sp.kron(sp.kron(I[0], I[1]), X[2])
    A = ksum(T).tocsc()
    B = (sp.identity(A.shape[0]) + eps * ksum(S)).tocsc()
    la, ls = [], []
    for n, (a, b, c) in zip(dims, coef):
        th = np.arange(1, n + 1) * np.pi / (n + 1)
        la.append(a + 2 * np.sqrt(b * c) * np.cos(th))
        ls.append(2 * np.sqrt(b * c) * np.cos(th))
    lamA = (la[0][:, None, None] + la[1][None, :, None] + la[2][None, None, :]).ravel()
    lamS = (ls[0][:, None, None] + ls[1][None, :, None] + ls[2][None, None, :]).ravel()
    lam = lamA / (1 + eps * lamS)
    Emid, rad = -2.064 + 0.2j, 0.5
    dist = np.abs(lam - Emid)
    inside = lam[dist <= rad]
    assert 4 <= len(inside) <= 20 and np.abs(dist - rad).min() > 1e-3
    n = A.shape[0]
    M0 = 32
    fpm = fc.feastinit()
    fpm[7] = 24
    fpm[2] = 10      # non-normal pencil: residuals level off near eps * cond(eigenvectors); the reference's general tests use 1e-7..1e-9
    Q0 = fo.seeded_subspace(n, M0)
    r = fc.pzifeast_gcsrgv(A, B, Emid, rad, M0, list(fpm), Q0=Q0, solver_tol=1e-12, solver_maxiter=4000, inner_rel=1e-10)
    assert r.info == 0 and r.M == len(inside)
    left = list(inside)
    for g in r.lambda_:
        j = int(np.argmin([abs(g - x) for x in left]))
        assert abs(g - left[j]) < 1e-9
        left.pop(j)
    assert r.res.max() < 1e-10
    ro = fo.feast_general(A, B, Emid, rad, M0, list(fpm), Q0=Q0, residual="true")
    assert ro.M == r.M


def _toeplitz_pencil(dims, eps=0.05):
    coef = [(0.4 + 0.1j, 1.0 + 0.05j, 0.9 - 0.05j), (0.3 - 0.1j, 0.8 + 0.1j, 0.75 + 0.05j), (0.5 + 0.2j, 0.6 - 0.05j, 0.65 + 0.02j)]

    def toe(n, a, b, c):
        return sp.diags([b * np.ones(n - 1), a * np.ones(n), c * np.ones(n - 1)], [-1, 0, 1])
    I = [sp.identity(n) for n in dims]
    T = [toe(n, *abc) for n, abc in zip(dims, coef)]
    S = [toe(n, 0.0, abc[1], abc[2]) for n, abc in zip(dims, coef)]
    ksum = lambda X: sp.kron(sp.kron(X[0], I[1]), I[2]) + sp.kron(sp.kron(I[0], X[1]), I[2]) + sp.kron(sp.kron(I[0], I[1]), X[2])
    A = ksum(T).tocsc()
    B = (sp.identity(A.shape[0]) + eps * ksum(S)).tocsc()
    la, ls = [], []
    for n, (a, b, c) in zip(dims, coef):
        th = np.arange(1, n + 1) * np.pi / (n + 1)
        la.append(a + 2 * np.sqrt(b * c) * np.cos(th))
        ls.append(2 * np.sqrt(b * c) * np.cos(th))
    lamA = (la[0][:, None, None] + la[1][None, :, None] + la[2][None, None, :]).ravel()
    lamS = (ls[0][:, None, None] + ls[1][None, :, None] + ls[2][None, None, :]).ravel()
    return A, B, lamA / (1 + eps * lamS)


def _disc_with(lam, count):
    """A disc at the left edge of the spectrum holding `count` eigenvalues (tools/run_config34.py's rule)."""
    order = np.argsort(lam.real)
    Emid = complex(lam[order[0]].real, lam[order[:count + 5]].imag.mean())
    dist = np.abs(lam - Emid)
    rad = 0.5 * (np.sort(dist)[count - 1] + np.sort(dist)[count])
    return Emid, rad, lam[dist <= rad]


def _check_general(r, inside, tol_exp):
    assert r.info == 0 and r.M == len(inside)
    left = list(inside)
    for g in r.lambda_:
        j = int(np.argmin([abs(g - x) for x in left]))
        assert abs(g - left[j]) < 1e-9
        left.pop(j)
    assert r.res.max() < 10.0 ** (-tol_exp)


def test_config3_fem_pair_on_the_generalized_lanczos_filter_108k():
    """configs[3] at 60x60x30 (n = 108 000, full size 100x100x50): zfeast_hcsrgv!, M0 = 96, 16 nodes, lowest 60 pairs, through the B-inner-
    product multi-shift Lanczos filter with Chebyshev inner solves (the per-node BiCGStab solves of round 1 stall at this size)."""
    import feastcuda as fc
    A, B, w = _fem_pair(60, 60, 30)
    n, M0, want = A.shape[0], 96, 60
    while w[want] - w[want - 1] < 1e-8 * w[want]:
        want += 1
    Emin, Emax = 0.0, 0.5 * (w[want - 1] + w[want])
    rng = np.random.default_rng(12345)
    Q0 = rng.standard_normal((n, M0)) + 0j
    Q0 /= np.linalg.norm(Q0, axis=0)
    fpm = fc.feastinit()
    fpm[1] = 16
    r = fc.zfeast_hcsrgv(A, B, Emin, Emax, M0, fpm, Q0=Q0, solver_maxiter=6000)
    assert r.info == 0 and r.M == want and r.loop <= 3
    assert np.abs(np.sort(r.lambda_) - w[:want]).max() < 1e-10 * w[want]
    assert r.res.max() < 1e-12 and r.stats["cheb_degree"] > 0 and r.stats["lz_steps_p1"] > 0
    R = A @ r.q - (B @ r.q) * r.lambda_
    assert (np.linalg.norm(R, axis=0) / np.maximum(np.abs(r.lambda_), 1.0)).max() < 1e-12     # independent residual check on the host
    G = r.q.conj().T @ (B @ r.q)
    D = np.sqrt(np.abs(np.diag(G)))
    offd = np.abs(G / np.outer(D, D) - np.eye(want)).max()
    assert offd < 1e-8                                                                          # B-orthogonal eigenvectors


@pytest.mark.skipif(not __import__("os").environ.get("FEASTCUDA_FULLSIZE"), reason="configs[3] at full size takes ~10 minutes: set FEASTCUDA_FULLSIZE=1")
def test_config3_full_size_500k():
    """configs[3] at FULL size (100x100x50, n = 500 000, M0 = 96, 16 nodes): info = 0, M = the analytic count, eigenvalues <= 1e-10 relative."""
    import feastcuda as fc
    A, B, w = _fem_pair(100, 100, 50)
    n, M0, want = A.shape[0], 96, 60
    while w[want] - w[want - 1] < 1e-8 * w[want]:
        want += 1
    Emin, Emax = 0.0, 0.5 * (w[want - 1] + w[want])
    rng = np.random.default_rng(12345)
    Q0 = rng.standard_normal((n, M0)) + 0j
    Q0 /= np.linalg.norm(Q0, axis=0)
    fpm = fc.feastinit()
    fpm[1] = 16
    r = fc.zfeast_hcsrgv(A, B, Emin, Emax, M0, fpm, Q0=Q0, solver_maxiter=6000)
    assert r.info == 0 and r.M == want
    assert np.abs(np.sort(r.lambda_) - w[:want]).max() < 1e-10 * w[want] and r.res.max() < 1e-12


def test_config4_general_pencil_two_sided_lanczos_16k():
    """configs[4] at 20x20x40 (n = 16 000): pzifeast_gcsrgv!, M0 = 64, 24 nodes, 35 eigenvalues in the disc, two-sided multi-shift Lanczos."""
    import feastcuda as fc
    A, B, lam = _toeplitz_pencil((20, 20, 40))
    Emid, rad, inside = _disc_with(lam, 35)
    assert len(inside) == 35
    n, M0 = A.shape[0], 64
    rng = np.random.default_rng(12345)
    Q0 = rng.standard_normal((n, M0)) + 0j
    Q0 /= np.linalg.norm(Q0, axis=0)
    fpm = fc.feastinit()
    fpm[7], fpm[2] = 24, 10       # non-normal pencil: the reference's general tests use 1e-7 .. 1e-9 (test/runtests.jl:204-222)
    r = fc.pzifeast_gcsrgv(A, B, Emid, rad, M0, fpm, Q0=Q0, solver_maxiter=4000)
    _check_general(r, inside, 10)
    assert r.stats["lz_steps_p1"] > 0


def test_config4_full_size_250k():
    """configs[4] at FULL size (50x50x100, n = 250 000, M0 = 64, 24 nodes, 35 eigenvalues in the disc): info = 0, the analytic eigenvalues
    to 1e-9, residuals below 10^-10 (~100 s on one B200; the reference's ne x M0 GMRES solves do not converge at this size)."""
    import feastcuda as fc
    A, B, lam = _toeplitz_pencil((50, 50, 100))
    Emid, rad, inside = _disc_with(lam, 35)
    assert len(inside) == 35
    n, M0 = A.shape[0], 64
    rng = np.random.default_rng(12345)
    Q0 = rng.standard_normal((n, M0)) + 0j
    Q0 /= np.linalg.norm(Q0, axis=0)
    fpm = fc.feastinit()
    fpm[7], fpm[2], fpm[3] = 24, 10, 30
    r = fc.pzifeast_gcsrgv(A, B, Emid, rad, M0, fpm, Q0=Q0, solver_maxiter=4000)
    _check_general(r, inside, 10)
    R = A @ r.q - (B @ r.q) * r.lambda_
    assert (np.linalg.norm(R, axis=0) / np.maximum(np.abs(r.lambda_), 1.0)).max() < 1e-10        # independent residual check on the host
