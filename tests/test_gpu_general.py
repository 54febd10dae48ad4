"""GPU parity tests of the general (non-Hermitian) solve: feast_general / feast_gcsrgv / feast_gegv / feast_gbgv
(kernel/feast_kernel.jl:646-962 behind the storage drivers).  Eigenvalues against the oracle and LAPACK, true residuals
||A x - lambda B x|| / max(|lambda|,1) below 10^-fpm[3]."""
import numpy as np
import pytest
import scipy.linalg as sla
import scipy.sparse as sp

import feast_oracle as fo

pytestmark = pytest.mark.gpu


def _match(got, want, tol):
    got, want = list(got), list(want)
    assert len(got) == len(want)
    for g in got:
        j = int(np.argmin([abs(g - w) for w in want]))
        assert abs(g - want[j]) <= tol, (g, want[j])
        want.pop(j)


def test_ka4_general_dense_sparse_standard_and_generalized():
    """runtests.jl:204-238 -- [1 2+i; 0 3], center 2, radius 2.5; B = diag(1,2)."""
    import feastcuda as fc
    A = np.array([[1, 2 + 1j], [0, 3]], dtype=complex)
    B = np.diag([1.0, 2.0]).astype(complex)
    Q0 = fo.seeded_subspace(2, 2)
    r = fc.feast_general(A, 2 + 0j, 2.5, M0=2, fpm=fc.feastinit(), Q0=Q0)
    assert r.info == 0 and r.M == 2
    _match(r.lambda_, np.linalg.eigvals(A), 1e-9)
    rs = fc.feast_general(sp.csc_matrix(A), 2 + 0j, 2.5, M0=2, fpm=fc.feastinit(), Q0=Q0)
    assert rs.info == 0 and rs.M == 2
    _match(rs.lambda_, np.linalg.eigvals(A), 1e-9)
    rg = fc.feast_general(A, B, 2 + 0j, 2.5, M0=2, fpm=fc.feastinit(), Q0=Q0)
    assert rg.info == 0 and rg.M == 2
    _match(rg.lambda_, sla.eigvals(A, B), 1e-9)
    rr = fc.feast_general(np.array([[1.0, 2.0], [0.0, 3.0]]), 2 + 0j, 2.5, M0=2, fpm=fc.feastinit(), Q0=Q0)   # real input is promoted
    assert rr.M == 2
    _match(rr.lambda_, [1.0, 3.0], 1e-9)
    with pytest.raises(ValueError):
        fc.feast_general(A, 2 + 0j, -1.0, M0=2, fpm=fc.feastinit())


def _toeplitz_pencil(n, seed):
    """Non-symmetric complex tridiagonal Toeplitz pencil with analytic eigenvalues a + 2 sqrt(bc) cos(k pi/(n+1))."""
    a, b, c = 0.3 + 0.2j, 1.0 + 0.1j, 0.95 - 0.05j   # mildly non-normal: eigenvector condition ~ |b/c|^(n/2) ~ 30
    A = sp.diags([b * np.ones(n - 1), a * np.ones(n), c * np.ones(n - 1)], [-1, 0, 1]).tocsc()
    k = np.arange(1, n + 1)
    lam = a + 2 * np.sqrt(b * c) * np.cos(k * np.pi / (n + 1))
    return A, lam


@pytest.mark.parametrize("kind", ["dense", "sparse", "band"])
def test_general_toeplitz_matches_analytic_and_oracle(kind):
    import feastcuda as fc
    n, M0 = 120, 24
    A, lam = _toeplitz_pencil(n, 0)
    Emid, r = 0.3 + 0.2j, 0.3
    inside = [l for l in lam if abs(l - Emid) <= r]
    assert 4 <= len(inside) <= 16
    Q0 = fo.seeded_subspace(n, M0)
    fpm = fc.feastinit()
    if kind == "dense":
        res = fc.feast_geev(A.toarray(), Emid, r, M0, fpm, Q0=Q0)
    elif kind == "sparse":
        res = fc.zifeast_gcsrev(A, Emid, r, M0, fpm, Q0=Q0, solver_tol=1e-13, solver_maxiter=4000)
    else:
        Af = A.toarray()
        AB = np.zeros((3, n), dtype=complex)
        for j in range(n):
            for i in range(max(0, j - 1), min(n, j + 2)):
                AB[1 + i - j, j] = Af[i, j]
        res = fc.feast_gbev(AB, 1, Emid, r, M0, fpm, Q0=Q0)
    assert res.info == 0 and res.M == len(inside)
    _match(res.lambda_, inside, 1e-9)
    assert res.res.max() < 1e-12
    Af = A.toarray()
    for j in range(res.M):                      # returned vectors are unit 2-norm eigenvectors
        x = res.q[:, j]
        assert abs(np.linalg.norm(x) - 1.0) < 1e-12
        assert np.linalg.norm(Af @ x - res.lambda_[j] * x) < 1e-11
    assert np.all(np.diff(np.abs(res.lambda_)) >= -1e-12)    # feast_sort_general!: ascending |lambda|
    ro = fo.feast_general(Af, None, Emid, r, M0, fo.feastinit(), Q0=Q0)
    assert ro.M == res.M
    _match(res.lambda_, ro.lambda_, 1e-9)


def test_general_generalized_dense_random():
    import feastcuda as fc
    n, M0 = 90, 30
    rng = np.random.default_rng(4)
    A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    B = np.eye(n) + 0.1 * (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    w = sla.eigvals(A, B)
    Emid, r = 1.0 + 1.0j, 2.2
    inside = [l for l in w if abs(l - Emid) <= r]
    assert 3 <= len(inside) <= 20
    edge = min(abs(abs(l - Emid) - r) for l in w)
    assert edge > 1e-3   # no eigenvalue sits on the contour
    res = fc.feast_gegv(A, B, Emid, r, M0, fc.feastinit(), Q0=fo.seeded_subspace(n, M0))
    assert res.info == 0 and res.M == len(inside)
    _match(res.lambda_, inside, 1e-8)
    for j in range(res.M):
        x = res.q[:, j]
        assert np.linalg.norm(A @ x - res.lambda_[j] * (B @ x)) / max(abs(res.lambda_[j]), 1.0) < 1e-10


def test_complex_symmetric_families_dense_sparse_banded():
    """runtests.jl:241-274 -- complex-symmetric 4x4 pencil (transpose-symmetric, not Hermitian), center 1+0.1i, radius 1.5:
    feast_gegv_complex_sym!/feast_geev_complex_sym! find eigvals(A, B) inside the contour (atol 1e-7 in the reference), reject a
    non-symmetric matrix; the sparse (feast_scsr*_complex!) and banded (feast_sb*_complex!) names reach the same values."""
    import feastcuda as fc
    import json
    from pathlib import Path
    k = json.loads((Path(__file__).parent / "golden" / "reference_known_answers.json").read_text())["KA15_complex_symmetric"]
    un = lambda c: np.array(c["re"]) + 1j * np.array(c["im"])
    A, B = un(k["A"]), np.diag(np.array(k["B_diag"], dtype=complex))
    Emid, r = complex(*k["center"]), k["radius"]
    want_g, want_s = list(un(k["expected_generalized"])), list(un(k["expected_standard"]))
    Q0 = fo.seeded_subspace(4, 4)
    rg = fc.feast_gegv_complex_sym(A.copy(), B.copy(), Emid, r, 4, fc.feastinit(), Q0=Q0)
    assert rg.info == 0 and rg.M == len(want_g)
    _match(rg.lambda_, want_g, 1e-9)
    rs = fc.feast_geev_complex_sym(A.copy(), Emid, r, 4, fc.feastinit(), Q0=Q0)
    assert rs.info == 0 and rs.M == len(want_s)
    _match(rs.lambda_, want_s, 1e-9)
    for j in range(rs.M):
        x = rs.q[:, j]
        assert np.linalg.norm(A @ x - rs.lambda_[j] * x) < 1e-10 * max(1.0, abs(rs.lambda_[j]))
    with pytest.raises(ValueError):
        fc.feast_geev_complex_sym(np.array([[1, 2], [0, 3]], dtype=complex), Emid, r, 2, fc.feastinit())
    rsp = fc.feast_scsrgv_complex(sp.csc_matrix(A), sp.csc_matrix(B), Emid, r, 4, fc.feastinit(), Q0=Q0)
    assert rsp.info == 0
    _match(rsp.lambda_, want_g, 1e-8)
    rse = fc.feast_scsrev_complex(sp.csc_matrix(A), Emid, r, 4, fc.feastinit(), Q0=Q0)
    _match(rse.lambda_, want_s, 1e-8)
    with pytest.raises(ValueError):
        fc.feast_scsrev_complex(sp.csc_matrix(np.array([[1, 2], [0, 3]], dtype=complex)), Emid, r, 2, fc.feastinit())
    rb = fc.feast_sbgv_complex(fo.full_to_banded(A, 1), fo.full_to_banded(B, 0), 1, 0, Emid, r, 4, fc.feastinit(), Q0=Q0)
    assert rb.info == 0
    _match(rb.lambda_, want_g, 1e-9)
    rbe = fc.feast_sbev_complex(fo.full_to_banded(A, 1), 1, Emid, r, 4, fc.feastinit(), Q0=Q0)
    _match(rbe.lambda_, want_s, 1e-9)


def test_polynomial_families_via_companion_linearisation():
    """feast_pep!/feast_gepev!/feast_sypev!/feast_polynomial (dense/feast_dense.jl:715-772,946-978; interfaces:448-462): a damped
    quadratic problem (lambda^2 M + lambda C + K) q = 0; eigenvalues inside the circle against LAPACK on the companion pencil,
    and P(lambda) q = 0 for the returned leading-block vectors."""
    import feastcuda as fc
    rng = np.random.default_rng(5)
    N = 6
    K = np.diag(np.linspace(1.0, 6.0, N)) + 0.05 * rng.standard_normal((N, N))
    K = 0.5 * (K + K.T)
    Cd = 0.1 * np.eye(N)
    M = np.eye(N)
    Al, Bl = fc.companion_linearization([K, Cd, M])
    w = sla.eigvals(Al, Bl)
    Emid, r = 0.0 + 1.6j, 0.75
    fpm0 = fo.feastdefault(fo.feastinit())
    want = [x for x in w if fo.feast_inside_gcontour(x, Emid, r, fpm0)]
    assert len(want) == 5                                   # M0 * d = 6 search columns
    res = fc.feast_sypev([K, Cd, M], 2, Emid, r, 3, fc.feastinit())
    assert res.info == 0 and res.M == len(want) and res.q.shape == (N, res.M)
    _match(res.lambda_, want, 1e-8)
    for j in range(res.M):
        lam, x = res.lambda_[j], res.q[:, j]
        assert np.linalg.norm((K + lam * Cd + lam * lam * M) @ x) < 1e-8 * np.linalg.norm(x)
    rp = fc.feast_polynomial([K.astype(complex), Cd.astype(complex), M.astype(complex)], Emid, r, M0=3, fpm=fc.feastinit())
    _match(rp.lambda_, want, 1e-8)
    with pytest.raises(ValueError):
        fc.feast_pep([K, Cd], 2, Emid, r, 3, fc.feastinit())
