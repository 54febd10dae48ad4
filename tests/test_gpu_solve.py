"""GPU parity tests of the whole FEAST solve through the reference-named API (C ABI underneath).

Bar (BASELINE.json north_star): same eigenvalue count M as the oracle, eigenvalues within 1e-10 relative,
every residual below 10^-fpm[3], eigenvector subspace angle below 1e-8.  The oracle is fed the SAME initial
subspace (Julia's seeded RNG is not reproducible outside Julia).
"""
import numpy as np
import pytest
import scipy.linalg as sla
import scipy.sparse as sp

import feast_oracle as fo

pytestmark = pytest.mark.gpu

# near-exact inner solves: 9 digits relative to the Ritz-guess residual, floor 1e-12 (BiCGStab's attainable accuracy)
TIGHT = dict(solver_tol=1e-12, solver_maxiter=4000, ritz_guess=True, inner_rel=1e-9)


def _check_pairs(r, ro, A, B=None, tol_exp=12, angle=1e-8):
    assert r.info == ro.info == 0
    assert r.M == ro.M
    lo, lg = np.sort(ro.lambda_), np.sort(r.lambda_)
    assert np.abs(lg - lo).max() <= 1e-10 * max(1.0, np.abs(lo).max())
    assert r.res.max() < 10.0 ** (-tol_exp)
    assert fo.subspace_angle(np.asarray(r.q, dtype=complex), np.asarray(ro.q, dtype=complex)) < angle


def test_ka1_tridiagonal_n3_all_eigenvalues():
    """runtests.jl:152-163 -- 1-D Laplacian n=3, (0.5,3.5), M0=3 (sparse route)."""
    import feastcuda as fc
    A = fo.laplacian_1d(3).tocsc()
    Q0 = fo.seeded_subspace(3, 3, complex_storage=False)
    r = fc.feast_scsrev(A, 0.5, 3.5, 3, fc.feastinit(), Q0=Q0, **TIGHT)
    assert r.info == 0 and r.M == 3
    assert np.allclose(np.sort(r.lambda_), np.linalg.eigvalsh(A.toarray()), atol=1e-10)
    assert r.q.dtype == np.float64  # FeastResult{T,T}: real.(q), dense/feast_dense.jl:372-387


def test_ka3_sparse_hermitian_tridiagonal():
    """runtests.jl:187-194 -- sparse complex Hermitian 3x3, (1.5,4.5)."""
    import feastcuda as fc
    v = np.array([0.1 + 0.2j, -0.05 + 0.1j])
    A = sp.diags([np.conj(v), np.array([2.0, 3.0, 4.0], dtype=complex), v], [-1, 0, 1]).tocsc()
    Q0 = fo.seeded_subspace(3, 3)
    r = fc.feast(A, (1.5, 4.5), M0=3, fpm=fc.feastinit(), Q0=Q0, **TIGHT)
    assert r.info == 0 and r.M == 3
    assert np.allclose(np.sort(r.lambda_), np.linalg.eigvalsh(A.toarray()), atol=1e-9)


def test_ka5_sparse_laplacian_n10_full_spectrum():
    """runtests.jl:399-412 -- n=10, [0,4], M0=10 -> M=10."""
    import feastcuda as fc
    A = fo.laplacian_1d(10).tocsc()
    r = fc.feast_scsrev(A, 0.0, 4.0, 10, fc.feastinit(), Q0=fo.seeded_subspace(10, 10, complex_storage=False), **TIGHT)
    assert r.info == 0 and r.M == 10
    assert np.allclose(np.sort(r.lambda_), np.linalg.eigvalsh(A.toarray()), atol=1e-10)


def test_ka6_diagonal_pencil_generalized_hermitian_and_custom_contour():
    """runtests.jl:416-440 -- A=diag(1..6), B=diag(1,1.2,1.5,2.5,4,5), [0.5,3.1]; x-variant agrees."""
    import feastcuda as fc
    A = sp.diags(np.arange(1.0, 7.0).astype(complex)).tocsc()
    B = sp.diags(np.array([1.0, 1.2, 1.5, 2.5, 4.0, 5.0], dtype=complex)).tocsc()
    expected = sorted(l for l in np.arange(1.0, 7.0) / np.array([1.0, 1.2, 1.5, 2.5, 4.0, 5.0]) if 0.5 <= l <= 3.1)
    Q0 = fo.seeded_subspace(6, 6)
    r = fc.feast_hcsrgv(A, B, 0.5, 3.1, 6, fc.feastinit(), Q0=Q0, **TIGHT)
    assert r.info == 0 and r.M == len(expected)
    assert np.allclose(np.sort(r.lambda_), expected, atol=1e-8)
    fpm = fc.feastinit()
    Z, W = fc.feast_contour(0.5, 3.1, fpm)
    rx = fc.feast_hcsrgvx(A, B, 0.5, 3.1, 6, fc.feastinit(), Z, W, Q0=Q0, **TIGHT)
    assert rx.info == 0 and np.allclose(np.sort(rx.lambda_), np.sort(r.lambda_), atol=1e-8)
    ro = fo.feast_hcsrgv(A, B, 0.5, 3.1, 6, fo.feastinit(), Q0=Q0)
    _check_pairs(r, ro, A, B)


def test_ka10_diag_parallel_alias():
    """runtests.jl:1042-1113 -- diag(0.5,1,1.5,3), [0.4,1.6] -> M=3 through the pd alias."""
    import feastcuda as fc
    A = sp.diags([0.5, 1.0, 1.5, 3.0]).tocsc()
    r = fc.pdfeast_scsrev(A, 0.4, 1.6, 4, fc.feastinit(), Q0=fo.seeded_subspace(4, 4, complex_storage=False), **TIGHT)
    assert r.info == 0 and r.M == 3
    assert np.allclose(np.sort(r.lambda_), [0.5, 1.0, 1.5], atol=1e-8)
    r2 = fc.dfeast_scsrev(A, 0.4, 1.6, 4, fc.feastinit(), Q0=fo.seeded_subspace(4, 4, complex_storage=False), **TIGHT)
    assert np.allclose(np.sort(r.lambda_), np.sort(r2.lambda_), atol=1e-10)  # alias == generic, runtests.jl:889-962


def test_ka11_oversized_subspace_rank_compression():
    """test_allocation_helpers.jl:319-346 -- n=80 diag(1..80), [10.5,12.5], M0=32, fpm[3]=7, fpm[4]=4 -> M=2."""
    import feastcuda as fc
    A = sp.diags(np.arange(1.0, 81.0)).tocsc()
    fpm = fc.feastinit()
    fpm[0], fpm[1], fpm[2], fpm[3] = 0, 8, 7, 4
    Q0 = fo.seeded_subspace(80, 32, complex_storage=False)
    for filt in ("reference", "true"):
        r = fc.feast_scsrev(A, 10.5, 12.5, 32, list(fpm), Q0=Q0, filter=filt, mixed=False, **TIGHT)   # FP64 filter: loop counts are compared
        ro = fo.feast_scsrev(A, 10.5, 12.5, 32, list(fpm), Q0=Q0.astype(complex), filter=filt)
        assert r.info == 0 and r.M == ro.M == 2
        assert np.allclose(np.sort(r.lambda_), [11.0, 12.0], atol=1e-8)
        assert r.res.max() < 1e-7
        assert r.loop == ro.loop  # same filter, same basis -> same number of refinement loops


@pytest.mark.parametrize("filt,N,M0,solver", [("reference", 10, 48, "bicgstab"), ("true", 14, 24, "bicgstab"),
                                              ("true", 14, 24, "mslanczos")])
def test_laplacian3d_matches_oracle_and_analytic(filt, N, M0, solver):
    """Reduced-size config 3: 7-point Laplacian N^3, lowest 10 eigenvalues (multiplicities 1,3,3,3).

    The reference's complex half-contour filter |g| decays like 1/distance; with M0=24 on 14^3
    the reference itself stops with M=0/info=5 at loop 0, so its parity case uses a wider subspace and more loops.
    """
    import feastcuda as fc
    A = fo.laplacian_3d(N).astype(float).tocsc()
    ev = fo.laplacian_3d_eigs(N)
    Emin, Emax = 0.0, 0.5 * (ev[9] + ev[10])
    Q0 = fo.seeded_subspace(N ** 3, M0, complex_storage=False)
    fpm = fc.feastinit()
    if filt == "reference":
        fpm[3] = 60
    r = fc.feast_scsrev(A, Emin, Emax, M0, list(fpm), Q0=Q0, filter=filt, solver=solver, **TIGHT)
    ro = fo.feast_scsrev(A, Emin, Emax, M0, list(fpm), Q0=Q0.astype(complex), filter=filt)
    _check_pairs(r, ro, A)
    assert (r.stats["lz_steps_p1"] > 0) == (solver == "mslanczos")
    assert np.abs(np.sort(r.lambda_) - ev[:10]).max() < 1e-10
    assert abs(r.loop - ro.loop) <= (1 if filt == "reference" else 0)


def test_reference_filter_failure_mode_is_reproduced():
    """Where the reference returns M=0 -> info=5 at loop 0 (weak complex filter), so does the engine in reference mode
    when its inner solves are exact enough."""
    import feastcuda as fc
    N, M0 = 14, 24
    A = fo.laplacian_3d(N).astype(float).tocsc()
    ev = fo.laplacian_3d_eigs(N)
    Emin, Emax = 0.0, 0.5 * (ev[9] + ev[10])
    Q0 = fo.seeded_subspace(N ** 3, M0, complex_storage=False)
    ro = fo.feast_scsrev(A, Emin, Emax, M0, fo.feastinit(), Q0=Q0.astype(complex), filter="reference")
    r = fc.feast_scsrev(A, Emin, Emax, M0, fc.feastinit(), Q0=Q0, filter="reference", **TIGHT)
    assert (ro.M, ro.info, ro.loop) == (0, 5, 0)
    assert (r.M, r.info, r.loop) == (0, 5, 0)


def test_inexact_inner_solves_with_ritz_guess_reach_full_accuracy():
    """The engine's bench setting: inner_rel=0.1, <=40 iterations per node per loop, Ritz-pair initial guess."""
    import feastcuda as fc
    N = 16
    A = fo.laplacian_3d(N).astype(float).tocsc()
    ev = fo.laplacian_3d_eigs(N)
    Emin, Emax = 0.0, 0.5 * (ev[9] + ev[10])
    M0 = 24
    Q0 = fo.seeded_subspace(N ** 3, M0, complex_storage=False)
    fpm = fc.feastinit()
    fpm[3] = 40
    r = fc.feast_scsrev(A, Emin, Emax, M0, fpm, Q0=Q0, filter="true", inner_rel=0.1, ritz_guess=True, solver_tol=1e-13,
                        solver_maxiter=40, solver_restart=0, solver="bicgstab")
    assert r.info == 0 and r.M == 10
    assert np.abs(np.sort(r.lambda_) - ev[:10]).max() < 1e-10
    assert r.res.max() < 1e-12
    # analytic eigenvectors: sine products; compare the invariant subspace of the 10 lowest eigenvalues
    ro = fo.feast_scsrev(A, Emin, Emax, M0, fo.feastinit(), Q0=Q0.astype(complex), filter="true")
    assert fo.subspace_angle(r.q.astype(complex), ro.q.astype(complex)) < 1e-8


def test_generalized_real_symmetric_fem_pair():
    """Reduced-size analogue of config 4's pencil in real arithmetic: K = kron sums, Mass = kron products."""
    import feastcuda as fc
    n1 = 9
    h = 1.0 / (n1 + 1)
    K1 = sp.diags([-np.ones(n1 - 1), 2 * np.ones(n1), -np.ones(n1 - 1)], [-1, 0, 1]) / h
    M1 = h * sp.diags([np.ones(n1 - 1), 4 * np.ones(n1), np.ones(n1 - 1)], [-1, 0, 1]) / 6
    K = (sp.kron(sp.kron(K1, M1), M1) + sp.kron(sp.kron(M1, K1), M1) + sp.kron(sp.kron(M1, M1), K1)).tocsc()
    Ms = sp.kron(sp.kron(M1, M1), M1).tocsc()
    import scipy.linalg as sla
    w = sla.eigh(K.toarray(), Ms.toarray(), eigvals_only=True)
    Emin, Emax = 0.0, 0.5 * (w[6] + w[7])
    M0 = 16
    Q0 = fo.seeded_subspace(K.shape[0], M0, complex_storage=False)
    r = fc.feast_scsrgv(K, Ms, Emin, Emax, M0, fc.feastinit(), Q0=Q0, filter="true", **TIGHT)
    ro = fo.feast_scsrgv(K, Ms, Emin, Emax, M0, fo.feastinit(), Q0=Q0.astype(complex), filter="true")
    _check_pairs(r, ro, K, Ms)
    assert np.abs(np.sort(r.lambda_) - w[:7]).max() < 1e-10 * w[6]


def test_empty_interval_returns_info5_and_argument_errors():
    import feastcuda as fc
    A = fo.laplacian_1d(10).tocsc()
    r = fc.feast_scsrev(A, 10.0, 11.0, 4, fc.feastinit(), Q0=fo.seeded_subspace(10, 4, complex_storage=False), **TIGHT)
    assert r.M == 0 and r.info == 5  # M == 0 -> Feast_ERROR_NO_CONVERGENCE (dense/feast_dense.jl:295-298)
    with pytest.raises(ValueError):
        fc.feast_scsrev(A, 1.0, 0.0, 4, fc.feastinit())
    with pytest.raises(ValueError):
        fc.feast_scsrev(A, 0.0, 1.0, 11, fc.feastinit())
    with pytest.raises(ValueError):
        fc.feast(sp.csc_matrix(np.array([[1.0, 2.0], [0.0, 3.0]])), (0.0, 4.0), M0=2)


def test_resident_three_call_form_equals_single_call(engine):
    import feastcuda as fc
    N = 10
    A = fo.laplacian_3d(N).astype(float).tocsc()
    ev = fo.laplacian_3d_eigs(N)
    Emin, Emax = 0.0, 0.5 * (ev[3] + ev[4])
    M0 = 12
    Q0 = fo.seeded_subspace(N ** 3, M0, complex_storage=False)
    engine.set_sparse(fc.A, A, fc.SYM)
    engine.clear_b()
    fpm = fc.feastinit()
    fc.feastdefault_(fpm)
    Z, W = fc.feast_contour(Emin, Emax, fpm)
    opts = engine.make_opts(filter="true", **TIGHT)
    engine.upload_subspace(M0, Q0)
    M, info, eps, loop = engine.run_interval(Emin, Emax, M0, fpm, Z, W, opts)
    lam, X, res = engine.fetch_results(M0, M, True)
    r = engine.solve_interval(Emin, Emax, M0, fc.feastinit(), Z, W, Q0=Q0, x_real=True, filter="true", **TIGHT)
    assert M == r.M == 4 and info == r.info == 0
    assert np.array_equal(lam, r.lambda_) and np.array_equal(X, r.q)


def test_mslanczos_inexact_matches_cpu_port_loop_for_loop():
    """The engine's default for real-symmetric standard problems (inner_rel=1e-3, Ritz guess) against its NumPy port
    (oracle/feast_port.py:feast_hrr_mslanczos): same loops, (nearly) the same Lanczos step counts, same eigenpairs."""
    import feastcuda as fc
    import feast_port as fp
    N, M0 = 16, 24
    A = fo.laplacian_3d(N).astype(float).tocsc()
    ev = fo.laplacian_3d_eigs(N)
    Emin, Emax = 0.0, 0.5 * (ev[9] + ev[10])
    Q0 = fo.seeded_subspace(N ** 3, M0, complex_storage=False)
    r = fc.feast_scsrev(A, Emin, Emax, M0, fc.feastinit(), Q0=Q0, solver_maxiter=2000, adaptive=False, mixed=False)   # FP64 recurrence = the port's
    rp = fp.feast_hrr_mslanczos(A.tocsr(), Emin, Emax, M0, fo.feastinit(), Q0, inner_rel=1e-3, inner_maxiter=2000, adaptive=False)
    ro = fo.feast_scsrev(A, Emin, Emax, M0, fo.feastinit(), Q0=Q0.astype(complex), filter="true")
    _check_pairs(r, ro, A)
    assert r.loop == rp.loop
    steps_port = sum(rp.stats["lz_steps"])
    assert abs(r.stats["lz_steps_p1"] - steps_port) <= max(4, 0.05 * steps_port)
    assert r.stats["lz_steps_p2"] == r.stats["lz_steps_p1"]
    assert fo.subspace_angle(r.q.astype(complex), rp.q.astype(complex)) < 1e-8
    # the adaptive last-sweep target (API default) reaches the same pairs in no more loops
    ra = fc.feast_scsrev(A, Emin, Emax, M0, fc.feastinit(), Q0=Q0, solver_maxiter=2000)
    _check_pairs(ra, ro, A)
    assert ra.loop <= r.loop


@pytest.mark.parametrize("N,M0,k_in", [(16, 24, 9), (20, 7, 3), (12, 61, 18)])
def test_mixed_precision_lanczos_reaches_fp64_pairs(N, M0, k_in):
    """fpm[42] = 1 ("single-precision solver", core/feast_parameters.jl:316-319; opts.mixed): FP32 Lanczos vectors inside the FP64
    refinement loop.  Same M, eigenvalues, residuals and subspace as the oracle; step counts track the NumPy port of the mixed
    recurrence (oracle/feast_port.py, mixed=True); column counts that are not multiples of 4 exercise the padded elements."""
    import feastcuda as fc
    import feast_port as fp
    A = fo.laplacian_3d(N).astype(float).tocsc()
    ev = fo.laplacian_3d_eigs(N)
    gaps = [i for i in range(k_in, k_in + 30) if ev[i + 1] - ev[i] > 1e-3]
    Emin, Emax = 0.0, 0.5 * (ev[gaps[0]] + ev[gaps[0] + 1])
    Q0 = fo.seeded_subspace(N ** 3, M0, complex_storage=False)
    ro = fo.feast_scsrev(A, Emin, Emax, M0, fo.feastinit(), Q0=Q0.astype(complex), filter="true")
    r = fc.feast_scsrev(A, Emin, Emax, M0, fc.feastinit(), Q0=Q0, solver_maxiter=2000, mixed=True)
    _check_pairs(r, ro, A)
    assert r.stats["lz_steps_fp32"] > 0 and r.stats["lz_steps_fp32"] <= r.stats["lz_steps_p1"]
    rp = fp.feast_hrr_mslanczos(A.tocsr(), Emin, Emax, M0, fo.feastinit(), Q0, inner_rel=1e-3, inner_maxiter=2000, adaptive=True, mixed=True)
    steps_port = sum(rp.stats["lz_steps"])
    assert rp.info == 0 and abs(r.loop - rp.loop) <= 1
    assert abs(r.stats["lz_steps_p1"] - steps_port) <= max(16, 0.25 * steps_port)
    # "fpm" mode follows fpm[42]: default 1 -> FP32 vectors, 0 -> FP64 only
    fpm = fc.feastinit()
    fpm[41] = 0
    r0 = fc.feast_scsrev(A, Emin, Emax, M0, fpm, Q0=Q0, solver_maxiter=2000, mixed="fpm")
    r1 = fc.feast_scsrev(A, Emin, Emax, M0, fc.feastinit(), Q0=Q0, solver_maxiter=2000, mixed="fpm")
    assert r0.stats["lz_steps_fp32"] == 0 and r1.stats["lz_steps_fp32"] > 0
    _check_pairs(r0, ro, A)
    _check_pairs(r1, ro, A)


@pytest.mark.parametrize("M0", [1, 2, 7, 33, 64, 100, 128])
def test_mslanczos_column_counts(M0):
    """Every lane mapping of the Lanczos SpMM (column pairs: 1..64 per row, odd counts, two chunks per lane)."""
    import feastcuda as fc
    n = 400
    rng = np.random.default_rng(3)
    d = np.sort(rng.uniform(1.0, 50.0, n))
    T = sp.diags([0.3 * np.ones(n - 1), d, 0.3 * np.ones(n - 1)], [-1, 0, 1]).tocsc()
    w = np.linalg.eigvalsh(T.toarray())
    want = max(1, min(M0 // 2, 20))
    Emin, Emax = w[0] - 0.1, 0.5 * (w[want - 1] + w[want])
    Q0 = fo.seeded_subspace(n, M0, complex_storage=False)
    fpm = fc.feastinit()
    fpm[3] = 40
    r = fc.feast_scsrev(T, Emin, Emax, M0, fpm, Q0=Q0, solver_maxiter=1000)
    assert r.info == 0 and r.M == want
    assert np.abs(np.sort(r.lambda_) - w[:want]).max() < 1e-10 * max(1.0, abs(w[want]))
    assert r.res.max() < 1e-12


def test_mslanczos_krylov_exhaustion_tiny_matrices():
    """n <= number of Lanczos steps: the recurrence terminates (beta = 0), frozen columns stay exact."""
    import feastcuda as fc
    for n, iv in ((3, (0.5, 3.5)), (10, (0.0, 4.0)), (2, (0.0, 5.0))):
        A = fo.laplacian_1d(n).tocsc()
        r = fc.feast_scsrev(A, iv[0], iv[1], n, fc.feastinit(), Q0=fo.seeded_subspace(n, n, complex_storage=False))
        assert r.info == 0 and r.M == n
        assert np.allclose(np.sort(r.lambda_), np.linalg.eigvalsh(A.toarray()), atol=1e-10)
        assert r.stats["lz_steps_p1"] > 0


def test_float32_entry_points_types_and_tolerance():
    """runtests.jl:281-304, 901-919, 1091-1113: s/c names return Float32 data, info == 0, eigenvalues within 1e-4/1e-5."""
    import feastcuda as fc
    A = sp.diags([0.5, 1.0, 1.5, 3.0]).tocsc().astype(np.float32)
    r = fc.psfeast_scsrev(A, np.float32(0.4), np.float32(1.6), 4, fc.feastinit(), Q0=fo.seeded_subspace(4, 4, complex_storage=False))
    assert r.info == 0 and r.M == 3
    assert r.lambda_.dtype == np.float32 and r.q.dtype == np.float32 and r.res.dtype == np.float32
    assert np.allclose(np.sort(r.lambda_), [0.5, 1.0, 1.5], atol=1e-5)
    N = 12
    L3 = fo.laplacian_3d(N).astype(np.float32).tocsc()
    ev = fo.laplacian_3d_eigs(N)
    Emin, Emax = 0.0, 0.5 * (ev[9] + ev[10])
    rs = fc.sfeast_scsrev(L3, Emin, Emax, 20, fc.feastinit(), Q0=fo.seeded_subspace(N ** 3, 20, complex_storage=False))
    rd = fc.dfeast_scsrev(L3.astype(np.float64), Emin, Emax, 20, fc.feastinit(), Q0=fo.seeded_subspace(N ** 3, 20, complex_storage=False),
                          mixed=False)
    assert rs.info == 0 and rs.M == rd.M == 10
    assert np.abs(np.sort(rs.lambda_) - ev[:10]).max() <= 1e-4 * ev[9]
    assert rs.res.max() <= np.sqrt(np.finfo(np.float32).eps) and rs.loop <= rd.loop   # stops at the Float32 tolerance
    assert rs.stats["lz_steps_fp32"] == rs.stats["lz_steps_p1"] > 0 and rd.stats["lz_steps_fp32"] == 0   # s-names: FP32 Krylov vectors
    Ah = np.array([[2.5, 0.2 + 0.1j, 0.0], [0.2 - 0.1j, 3.5, 0.3 - 0.2j], [0.0, 0.3 + 0.2j, 4.0]], dtype=np.complex64)
    rc = fc.cfeast_heev(Ah, 2.0, 5.0, 3, fc.feastinit(), Q0=fo.seeded_subspace(3, 3))
    assert rc.info == 0 and rc.M == 3 and rc.q.dtype == np.complex64 and rc.lambda_.dtype == np.float32
    assert np.allclose(np.sort(rc.lambda_), np.linalg.eigvalsh(Ah.astype(np.complex128)), atol=1e-4)


@pytest.mark.parametrize("M0", [16, 40])
def test_complex_hermitian_standard_runs_the_lanczos_path(M0):
    """zfeast_hcsrev! (sparse/feast_sparse.jl:759-788) on a gauge-transformed 3-D Laplacian: complex Hermitian, analytic spectrum.
    The engine runs the multi-shift Lanczos recurrence on complex vectors (true filter, real tridiagonal)."""
    import feastcuda as fc
    import feast_port as fp
    N = 10
    n = N ** 3
    L = fo.laplacian_3d(N).astype(complex).tocsr()
    phi = np.random.default_rng(7).uniform(0, 2 * np.pi, n)
    D = sp.diags(np.exp(1j * phi))
    A = (D @ L @ D.conj()).tocsr()
    A = ((A + A.conj().T) * 0.5).tocsc()
    ev = fo.laplacian_3d_eigs(N)
    Emin, Emax = 0.0, 0.5 * (ev[9] + ev[10])
    Q0 = fo.seeded_subspace(n, M0)
    r = fc.zfeast_hcsrev(A, Emin, Emax, M0, fc.feastinit(), Q0=Q0, solver_maxiter=2000)
    ro = fo.feast_hcsrev(A, Emin, Emax, M0, fo.feastinit(), Q0=Q0, filter="true")
    _check_pairs(r, ro, A)
    assert r.stats["lz_steps_p1"] > 0 and r.q.dtype == np.complex128
    assert np.abs(np.sort(r.lambda_) - ev[:10]).max() < 1e-10
    rp = fp.feast_hrr_mslanczos(A.tocsr(), Emin, Emax, M0, fo.feastinit(), Q0, inner_rel=1e-3, inner_maxiter=2000, adaptive=True)
    assert abs(r.loop - rp.loop) <= 1
    assert fo.subspace_angle(r.q, rp.q) < 1e-8
    # the per-node complex BiCGStab (reference filter) reaches the same pairs
    fpm = fc.feastinit()
    fpm[3] = 60
    rb = fc.zfeast_hcsrev(A, Emin, Emax, 48, list(fpm), Q0=fo.seeded_subspace(n, 48), solver="bicgstab", filter="reference",
                          solver_tol=1e-12, solver_maxiter=4000, ritz_guess=True, inner_rel=1e-9)
    assert rb.info == 0 and rb.M == 10 and np.abs(np.sort(rb.lambda_) - ev[:10]).max() < 1e-10


def _rci_drive(fn, A, B, N, M0, Emin, Emax, engine, hermitian, maxiter=400, stop_at_first_mult=False, solver="bicgstab"):
    """The caller's side of the reverse-communication loop (banded/feast_banded.jl:76-180 shape): factorize = remember the
    shift, solve = one block solve on the device, mult_a = A q on the device."""
    import feastcuda as fc
    ijob, Ze, eps, loop, mode, info = fc.Ref(-1), fc.Ref(0j), fc.Ref(0.0), fc.Ref(0), fc.Ref(0), fc.Ref(-1)
    work = np.zeros((N, M0))
    workc = np.zeros((N, M0), dtype=complex)
    dt = complex if hermitian else float
    Aq, Sq = np.zeros((M0, M0), dtype=dt), np.zeros((M0, M0), dtype=dt)
    lam, q, res = np.zeros(M0), np.zeros((N, M0), dtype=dt), np.zeros(M0)
    fpm = fc.feastinit()
    state = fc.FeastRCIState()
    z = 0j
    for _ in range(maxiter):
        fn(ijob, N, Ze, work, workc, Aq, Sq, fpm, eps, loop, Emin, Emax, M0, lam, q, mode, res, info, state=state, engine=engine)
        if ijob.v == 10:
            z = Ze.v
        elif ijob.v == 11:
            rhs = workc[:, :M0] if hermitian else work[:, :M0].astype(complex)
            if B is not None:
                rhs = engine.apply(fc.B, rhs)
            X, _, _ = engine.block_solve(z, rhs, solver_tol=1e-13, solver_maxiter=4000, solver=solver)
            workc[:, :M0] = X
        elif ijob.v == 30:
            if stop_at_first_mult:
                break
            AX = engine.apply(fc.A, np.asarray(q[:, :mode.v], dtype=complex))
            if hermitian:
                workc[:, :mode.v] = AX
            else:
                work[:, :mode.v] = AX.real
        elif ijob.v == 0:
            break
    return lam[:mode.v].copy(), np.array(q[:, :mode.v]), res[:mode.v].copy(), info.v, loop.v, eps.v, state


def test_rci_state_machines_drive_a_full_solve(engine):
    """feast_srci! / feast_hrci! (kernel/feast_kernel.jl:7-293, 397-644) with the caller's work done through the stage-level
    entry points: KA10 (diag(0.5,1,1.5,3), [0.4,1.6] -> 0.5,1,1.5), the n=10 1-D Laplacian (test_parallel_backends.jl:63-168)
    and a 3x3 complex Hermitian matrix; compared with the oracle's moment solver."""
    import feastcuda as fc
    A = sp.diags([0.5, 1.0, 1.5, 3.0]).tocsc()
    engine.set_sparse(fc.A, A, fc.SYM)
    engine.clear_b()
    lam, q, res, info, loop, eps, st = _rci_drive(fc.feast_srci, A, None, 4, 4, 0.4, 1.6, engine, False)
    assert info == 0 and np.allclose(lam, [0.5, 1.0, 1.5], atol=1e-8) and res.max() < 1e-10
    L1 = fo.laplacian_1d(10).tocsc()
    # the reference drives feast_srci! with band LU factorisations (banded/feast_banded.jl:76-180): same here, on the device
    engine.set_band(fc.A, fo.full_to_banded(L1.toarray(), 1), 1, fc.SYM)
    engine.clear_b()
    lam, q, res, info, loop, eps, st = _rci_drive(fc.pdfeast_srci, L1, None, 10, 8, 0.1, 1.9, engine, False, solver="direct")
    w = np.linalg.eigvalsh(L1.toarray())
    want = w[(w >= 0.1) & (w <= 1.9)]
    # M0 = 8 > 4 eigenvalues inside: Aq is rank deficient; the null directions must be deflated, not returned as Ritz values
    # (LAPACK's QZ, hence the reference and the oracle, may report an arbitrary value for them)
    assert info == 0, st.extra
    assert len(lam) == len(want) and np.allclose(lam, want, atol=1e-8)
    lam5, _, res5, info5, _, _, st5 = _rci_drive(fc.feast_srci, L1, None, 10, 5, 0.1, 1.9, engine, False, solver="direct")
    ro = fo.feast_smom(L1, None, 0.1, 1.9, 5, fo.feastinit())
    assert info5 == ro.info == 0 and len(lam5) == ro.M and np.allclose(lam5, ro.lambda_[:ro.M], atol=1e-8), st5.extra
    assert np.all(np.diff(lam) >= 0) and res.max() < 1e-10                      # feast_sort!: ascending at exit
    for j in range(len(lam)):
        assert np.linalg.norm(L1 @ q[:, j] - lam[j] * q[:, j]) / np.linalg.norm(q[:, j]) < 1e-9
    # H-MOM keeps the reference's half-contour sums 2 w_e Y without a Hermitian part (kernel/feast_kernel.jl:516-524), so its
    # first-sweep Ritz values are those of (Q0^H h(A) Q0, Q0^H g(A) Q0), g = sum 2w/(z-x), h = sum 2wz/(z-x): restated here
    v = np.array([0.1 + 0.2j, -0.05 + 0.1j])
    Ah = sp.diags([np.conj(v), np.array([2.0, 3.0, 4.0], dtype=complex), v], [-1, 0, 1]).tocsc()
    engine.set_dense(fc.A, Ah.toarray(), fc.HERM)
    engine.clear_b()
    lam, q, res, info, loop, eps, st = _rci_drive(fc.feast_hrci, Ah, None, 3, 3, 1.5, 4.5, engine, True, stop_at_first_mult=True, solver="direct")
    Z, W = fo.feast_contour(1.5, 4.5, fo.feastdefault(fo.feastinit()))
    Ad = Ah.toarray()
    G = sum(2 * w * np.linalg.inv(z * np.eye(3) - Ad) for z, w in zip(Z, W))
    H = sum(2 * w * z * np.linalg.inv(z * np.eye(3) - Ad) for z, w in zip(Z, W))
    Q0 = st.Q0
    want_h = np.sort(sla.eigvals(Q0.conj().T @ H @ Q0, Q0.conj().T @ G @ Q0).real)
    assert np.allclose(st.zAq, Q0.conj().T @ G @ Q0, atol=1e-10) and np.allclose(st.zSq, Q0.conj().T @ H @ Q0, atol=1e-10)
    assert np.allclose(np.sort(lam), want_h[(want_h >= 1.5) & (want_h <= 4.5)], atol=1e-9)
    assert abs(lam[np.argmin(abs(lam - 3.0))] - np.linalg.eigvalsh(Ad)[1]) < 0.05          # exact only near the centre


def test_zolotarev_quadrature_solves_like_the_oracle():
    """fpm[16] = 2 (Zolotarev rule, core/feast_tools.jl:44-210,263-266): the same contour in libfeastcuda and in the oracle, same pairs."""
    import feastcuda as fc
    n = 60
    A = fo.laplacian_1d(n).tocsc()
    lam = 2 - 2 * np.cos(np.arange(1, n + 1) * np.pi / (n + 1))
    Emin, Emax = 0.5, 1.5
    inside = lam[(lam >= Emin) & (lam <= Emax)]
    M0 = 2 * len(inside)
    Q0 = fo.seeded_subspace(n, M0, complex_storage=False)
    fpm, fpo = fc.feastinit(), fo.feastinit()
    fpm[15] = fpo[15] = 2
    r = fc.feast_scsrev(A, Emin, Emax, M0, fpm, Q0=Q0, **TIGHT)
    ro = fo.feast_scsrev(A, Emin, Emax, M0, fpo, Q0=Q0.astype(complex), filter="true")
    assert r.info == ro.info == 0 and r.M == ro.M == len(inside)
    assert np.abs(np.sort(r.lambda_) - inside).max() < 1e-10 and r.res.max() < 1e-12
    assert fo.subspace_angle(r.q.astype(complex), ro.q.astype(complex)) < 1e-8
    rd = fc.feast_syev(A.toarray(), Emin, Emax, M0, list(fpm), Q0=Q0)          # dense route: the cached LU per Zolotarev node
    assert rd.info == 0 and rd.M == len(inside) and np.abs(np.sort(rd.lambda_) - inside).max() < 1e-10
