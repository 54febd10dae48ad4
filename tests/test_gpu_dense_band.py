"""GPU parity tests of the dense (batched LU on the FP64 tensor pipe) and banded (band LU) operator paths, through the
reference-named API.  Bar: same M as the oracle, eigenvalues within 1e-10 relative, residuals < 10^-fpm[3], subspace angle
< 1e-8 (north star)."""
import numpy as np
import pytest
import scipy.linalg as sla

import feast_oracle as fo

pytestmark = pytest.mark.gpu


def _check_pairs(r, ro, tol_exp=12, angle=1e-8):
    assert r.info == ro.info == 0
    assert r.M == ro.M
    lo, lg = np.sort(ro.lambda_), np.sort(r.lambda_)
    assert np.abs(lg - lo).max() <= 1e-10 * max(1.0, np.abs(lo).max())
    assert r.res.max() < 10.0 ** (-tol_exp)
    assert fo.subspace_angle(np.asarray(r.q, dtype=complex), np.asarray(ro.q, dtype=complex)) < angle


def _sym(n, seed, spread=10.0):
    rng = np.random.default_rng(seed)
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    d = np.sort(rng.uniform(0.0, spread, n))
    A = (Q * d) @ Q.T
    return 0.5 * (A + A.T), d


def test_ka1_ka2_ka10_dense_known_answers():
    """runtests.jl:152-178, 1042-1061."""
    import feastcuda as fc
    A = 2 * np.eye(3) - np.eye(3, k=1) - np.eye(3, k=-1)
    r = fc.feast(A, np.eye(3), (0.5, 3.5), M0=3, fpm=fc.feastinit(), Q0=fo.seeded_subspace(3, 3, complex_storage=False))
    assert r.info == 0 and r.M == 3 and np.allclose(np.sort(r.lambda_), np.linalg.eigvalsh(A), atol=1e-10)
    assert r.q.dtype == np.float64
    Ah = np.array([[2.5, 0.2 + 0.1j, 0.0], [0.2 - 0.1j, 3.5, 0.3 - 0.2j], [0.0, 0.3 + 0.2j, 4.0]])
    r = fc.feast(Ah, (2.0, 5.0), M0=3, fpm=fc.feastinit(), Q0=fo.seeded_subspace(3, 3))
    assert r.info == 0 and r.M == 3 and np.allclose(np.sort(r.lambda_), np.linalg.eigvalsh(Ah), atol=1e-9)
    r = fc.dfeast_syev(np.diag([0.5, 1.0, 1.5, 3.0]), 0.4, 1.6, 4, fc.feastinit(), Q0=fo.seeded_subspace(4, 4, complex_storage=False))
    assert r.info == 0 and r.M == 3 and np.allclose(np.sort(r.lambda_), [0.5, 1.0, 1.5], atol=1e-8)


def test_ka11_dense_and_banded_rank_compression():
    """test_allocation_helpers.jl:294-346: n=80 diag(1..80), [10.5,12.5], M0=32, fpm[3]=7, fpm[4]=4 -> M=2."""
    import feastcuda as fc
    n = 80
    d = np.arange(1.0, n + 1)
    fpm = fc.feastinit()
    fpm[0], fpm[1], fpm[2], fpm[3] = 0, 8, 7, 4
    Q0 = fo.seeded_subspace(n, 32, complex_storage=False)
    r = fc.feast_syev(np.diag(d), 10.5, 12.5, 32, list(fpm), Q0=Q0)
    assert r.info == 0 and r.M == 2 and np.allclose(np.sort(r.lambda_), [11.0, 12.0], atol=1e-8) and r.res.max() < 1e-7
    rb = fc.feast_hbev(d.reshape(1, n).astype(complex), 0, 10.5, 12.5, 32, list(fpm), Q0=Q0.astype(complex))
    assert rb.info == 0 and rb.M == 2 and np.allclose(np.sort(rb.lambda_), [11.0, 12.0], atol=1e-8) and rb.res.max() < 1e-7


@pytest.mark.parametrize("n,M0", [(150, 24), (333, 40), (700, 48)])
def test_dense_real_symmetric_matches_oracle(n, M0):
    """Reduced config 2: dense symmetric with known spectrum; sizes that are not multiples of the LU / GEMM tiles."""
    import feastcuda as fc
    A, d = _sym(n, n)
    want = 12
    Emin, Emax = 0.5 * (d[19] + d[20]), 0.5 * (d[19 + want] + d[20 + want])
    Q0 = fo.seeded_subspace(n, M0, complex_storage=False)
    r = fc.feast_syev(A, Emin, Emax, M0, fc.feastinit(), Q0=Q0)
    ro = fo.feast_syev(A, Emin, Emax, M0, fo.feastinit(), Q0=Q0.astype(complex), filter="true")
    _check_pairs(r, ro)
    assert r.loop == ro.loop
    assert np.abs(np.sort(r.lambda_) - d[20:20 + want]).max() < 1e-10 * d[-1]


def test_dense_reference_filter_and_generalized_hermitian():
    import feastcuda as fc
    n, M0 = 200, 30
    rng = np.random.default_rng(5)
    G = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    A = (G + G.conj().T) / 2
    Bh = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    B = Bh @ Bh.conj().T / n + 2 * np.eye(n)
    B = 0.5 * (B + B.conj().T)
    w = sla.eigh(A, B, eigvals_only=True)
    Emin, Emax = 0.5 * (w[89] + w[90]), 0.5 * (w[99] + w[100])
    Q0 = fo.seeded_subspace(n, M0)
    r = fc.feast_hegv(A, B, Emin, Emax, M0, fc.feastinit(), Q0=Q0, filter="reference")
    ro = fo.feast_hegv(A, B, Emin, Emax, M0, fo.feastinit(), Q0=Q0, filter="reference")
    _check_pairs(r, ro)
    assert np.abs(np.sort(r.lambda_) - w[90:100]).max() < 1e-10 * np.abs(w).max()
    As, d = _sym(160, 9)
    Bs = np.diag(np.linspace(1.0, 2.0, 160))
    ws = sla.eigh(As, Bs, eigvals_only=True)
    Q0r = fo.seeded_subspace(160, 20, complex_storage=False)
    r2 = fc.feast_sygv(As, Bs, 0.5 * (ws[29] + ws[30]), 0.5 * (ws[37] + ws[38]), 20, fc.feastinit(), Q0=Q0r)
    assert r2.info == 0 and r2.M == 8 and np.abs(np.sort(r2.lambda_) - ws[30:38]).max() < 1e-10 * ws[-1] and r2.res.max() < 1e-12


def test_dense_block_solve_stage_matches_lapack(engine):
    import feastcuda as fc
    n, m = 257, 37
    rng = np.random.default_rng(11)
    A, _ = _sym(n, 3)
    engine.set_dense(fc.A, A, fc.SYM)
    engine.clear_b()
    z = 4.3 + 0.21j
    RHS = rng.standard_normal((n, m)) + 1j * rng.standard_normal((n, m))
    X, _, _ = engine.block_solve(z, RHS, solver="direct")
    want = np.linalg.solve(z * np.eye(n) - A, RHS)
    assert np.abs(X - want).max() < 1e-11 * np.abs(want).max()
    Y = engine.apply(fc.A, RHS)
    assert np.abs(Y - A @ RHS).max() < 1e-12 * np.abs(A @ RHS).max()


def test_banded_known_answers_and_random_pencil(engine):
    """runtests.jl:605-636 (1-D Laplacian n=8 in band storage, (0.5, 3.1)) and a random symmetric band pencil."""
    import feastcuda as fc
    A = fo.laplacian_1d(8).toarray()
    AB = fo.full_to_banded(A, 1)
    w = np.linalg.eigvalsh(A)
    want = w[(w >= 0.5) & (w <= 3.1)]
    r = fc.feast_banded(AB, 1, (0.5, 3.1), M0=8, fpm=fc.feastinit(), Q0=fo.seeded_subspace(8, 8, complex_storage=False))
    assert r.info == 0 and r.M == len(want) and np.allclose(np.sort(r.lambda_), want, atol=1e-8)
    n, k, M0 = 600, 7, 24
    rng = np.random.default_rng(2)
    Af = np.zeros((n, n))
    Bf = np.zeros((n, n))
    for dd in range(k + 1):
        v = rng.standard_normal(n - dd) * (3.0 if dd == 0 else 0.4)
        Af += np.diag(v, dd) + (np.diag(v, -dd) if dd else 0)
        if dd <= 2:
            u = rng.uniform(0.05, 0.1, n - dd) if dd else rng.uniform(1.0, 2.0, n)
            Bf += np.diag(u, dd) + (np.diag(u, -dd) if dd else 0)
    ws = sla.eigh(Af, Bf, eigvals_only=True)
    gaps = np.diff(ws)
    lo = next(i for i in range(200, 400) if gaps[i - 1] > 0.02 and gaps[i + 9] > 0.02)   # well separated interval ends
    Emin, Emax = 0.5 * (ws[lo - 1] + ws[lo]), 0.5 * (ws[lo + 9] + ws[lo + 10])
    Q0 = fo.seeded_subspace(n, M0, complex_storage=False)
    r = fc.feast_sbgv(fo.full_to_banded(Af, k), fo.full_to_banded(Bf, 2), k, 2, Emin, Emax, M0, fc.feastinit(), Q0=Q0)
    assert r.info == 0 and r.M == 10
    assert np.abs(np.sort(r.lambda_) - ws[lo:lo + 10]).max() < 1e-10 * np.abs(ws).max() and r.res.max() < 1e-12
    ro = fo.feast_sygv(Af, Bf, Emin, Emax, M0, fo.feastinit(), Q0=Q0.astype(complex), filter="true")
    _check_pairs(r, ro)
    # Hermitian band, standard problem
    Ah = Af.astype(complex)
    ph = np.exp(1j * rng.uniform(0, 2 * np.pi, n))
    Ah = (ph[:, None] * Ah) * ph.conj()[None, :]
    Ah = 0.5 * (Ah + Ah.conj().T)
    wh = np.linalg.eigvalsh(Ah)
    gh = np.diff(wh)
    l2 = next(i for i in range(100, 400) if gh[i - 1] > 0.02 and gh[i + 7] > 0.02)
    # complex Hermitian pencils run the reference's complex half-contour filter (no adjoint solves): slower decay, more loops
    fpm_h = fc.feastinit()
    fpm_h[3] = 80
    Q0h = fo.seeded_subspace(n, 32)
    rh = fc.feast_hbev(fo.full_to_banded(Ah, k), k, 0.5 * (wh[l2 - 1] + wh[l2]), 0.5 * (wh[l2 + 7] + wh[l2 + 8]), 32, list(fpm_h),
                       Q0=Q0h, filter="reference")
    assert rh.info == 0 and rh.M == 8 and np.abs(np.sort(rh.lambda_) - wh[l2:l2 + 8]).max() < 1e-10 * np.abs(wh).max()
    roh = fo.feast_heev(Ah, 0.5 * (wh[l2 - 1] + wh[l2]), 0.5 * (wh[l2 + 7] + wh[l2 + 8]), 32, list(fpm_h), Q0=Q0h, filter="reference")
    _check_pairs(rh, roh)
    # stage-level band solve against a dense solve
    engine.set_band(fc.A, fo.full_to_banded(Af, k), k, fc.SYM)
    engine.clear_b()
    z = 0.3 + 0.05j
    RHS = rng.standard_normal((n, 9)) + 1j * rng.standard_normal((n, 9))
    X, _, _ = engine.block_solve(z, RHS, solver="direct")
    want_x = np.linalg.solve(z * np.eye(n) - Af, RHS)
    assert np.abs(X - want_x).max() < 1e-10 * np.abs(want_x).max()


def test_fixture_files_feed_the_sparse_dense_and_banded_drivers(tmp_path):
    """The on-disk input format of the original FEAST example systems (examples/feast/utils.jl:15-150) read by feastcuda.fixtures and
    solved by the reference-named drivers: the KA5 / KA8 operators (runtests.jl:399-412, 605-636) written to a file first."""
    import feastcuda as fc
    A = fo.laplacian_1d(10)
    fc.write_mm_coordinate(tmp_path / "lap10.mtx", A)
    w = np.linalg.eigvalsh(A.toarray())
    want = w[(w >= 0.5) & (w <= 3.1)]
    Q0 = fo.seeded_subspace(10, 8, complex_storage=False)
    rs = fc.feast_scsrev(fc.read_mm_sparse_real("lap10", data_dir=tmp_path), 0.5, 3.1, 8, fc.feastinit(), Q0=Q0, **{
        "solver_tol": 1e-12, "solver_maxiter": 4000, "ritz_guess": True, "inner_rel": 1e-9})
    rd = fc.feast_syev(fc.read_mm_dense_real("lap10", data_dir=tmp_path), 0.5, 3.1, 8, fc.feastinit(), Q0=Q0)
    band, kl, ku = fc.read_banded_real("lap10", data_dir=tmp_path)
    assert (kl, ku) == (1, 1)
    rb = fc.feast_sbev(np.ascontiguousarray(band[:ku + 1]), ku, 0.5, 3.1, 8, fc.feastinit(), Q0=Q0)   # upper band = rows 0..ku
    rg = fc.feast_gbev(band.astype(complex), kl, 1.8, 1.3, 8, fc.feastinit(), Q0=fo.seeded_subspace(10, 8))
    for r in (rs, rd, rb):
        assert r.info == 0 and r.M == len(want) and np.allclose(np.sort(r.lambda_), want, atol=1e-10) and r.res.max() < 1e-12
    assert rg.info == 0 and rg.M == len(want) and np.allclose(np.sort(rg.lambda_.real), want, atol=1e-8)


@pytest.mark.parametrize("n,k,m,generalized", [(5000, 7, 40, False), (777, 12, 33, True), (300, 20, 5, False), (200, 40, 64, False),
                                               (37, 1, 1, False), (64, 3, 70, True), (3, 1, 3, False), (1, 0, 2, False)])
def test_band_block_solve_every_kernel_path_matches_lapack(engine, n, k, m, generalized):
    """Stage-level band solve (feastcuda_block_solve on a banded pencil = zgbtrf + zgbtrs, banded/feast_banded.jl:108,141) against LAPACK
    for every kernel choice: the warp LU + register-window solve with its four window sizes (k <= 2, 4, 8, 16), the warp LU with the
    thread-per-column solve (16 < k <= 32) and the one-CTA LU (k > 32); column counts that do not fill a warp; the chunk tails of the
    shared-memory staging (n not a multiple of 16, n smaller than a window)."""
    import feastcuda as fc
    rng = np.random.default_rng(7 * n + k)
    Af = np.zeros((n, n))
    Bf = np.zeros((n, n))
    kb = min(k, 2)
    for dd in range(min(k, n - 1) + 1):
        v = rng.standard_normal(n - dd) * (2.0 if dd == 0 else 0.5)
        Af += np.diag(v, dd) + (np.diag(v, -dd) if dd else 0)
        if dd <= kb:
            u = rng.uniform(0.05, 0.1, n - dd) if dd else rng.uniform(1.0, 2.0, n)
            Bf += np.diag(u, dd) + (np.diag(u, -dd) if dd else 0)
    engine.set_band(fc.A, fo.full_to_banded(Af, k), k, fc.SYM)
    if generalized:
        engine.set_band(fc.B, fo.full_to_banded(Bf, kb), kb, fc.SYM)
    else:
        engine.clear_b()
    z = 0.3 + 0.05j
    S = z * (Bf if generalized else np.eye(n)) - Af
    RHS = rng.standard_normal((n, m)) + 1j * rng.standard_normal((n, m))
    want = sla.solve_banded((k, k), fo.full_to_general_banded(S, k), RHS) if n > 1 else RHS / S[0, 0]
    X, _, _ = engine.block_solve(z, RHS, solver="direct")
    assert np.abs(X - want).max() < 1e-10 * np.abs(want).max()
    # a second solve with the cached factor and other right-hand sides
    RHS2 = rng.standard_normal((n, m)) + 0j
    X2, _, _ = engine.block_solve(z, RHS2, solver="direct")
    want2 = sla.solve_banded((k, k), fo.full_to_general_banded(S, k), RHS2) if n > 1 else RHS2 / S[0, 0]
    assert np.abs(X2 - want2).max() < 1e-10 * np.abs(want2).max()
