"""CPU tests of the host-side API helpers (feastcuda/utils.py): the reference's "Performance utilities" test set
(test/runtests.jl:1224-1246) plus the rational-filter evaluators and the custom-contour weights."""
import io

import numpy as np
import pytest
import scipy.sparse as sp

import feast_oracle as fo


def test_performance_utilities_like_the_reference_testset():
    import feastcuda as fc
    buf = io.StringIO()
    assert fc.feast_memory_estimate(50, 8, np.float64, io=buf) == 50 * 8 * 8 + 50 * 8 * 16 + 2 * 64 * 8 + (50 * 8 + 16) * 8
    assert "Total estimate" in buf.getvalue()
    assert fc.feast_memory_estimate(10 ** 6, 64, device=True, io=io.StringIO()) == 12 * 10 ** 6 * 64 * 16
    A = np.diag([1.0, 2.0, 3.0, 4.0])
    lo, hi = fc.feast_validate_interval(A, (1.5, 3.5))
    assert lo <= hi and (lo, hi) == (1.0, 4.0)
    L1 = fo.laplacian_1d(6)
    assert fc.feast_validate_interval(sp.csc_matrix(L1), (0.5, 1.0)) == (0.0, 4.0)
    with pytest.raises(ValueError):
        fc.feast_validate_interval(A, (2.0, 1.0))
    with pytest.warns(UserWarning):
        fc.feast_validate_interval(A, (10.0, 11.0))
    res = fc.FeastResult(np.array([1.0, 2.0]), np.eye(2), 2, np.array([1e-12, 1e-12]), 0, 1e-12, 3)
    out = io.StringIO()
    fc.feast_summary(res, out)
    assert "Eigenvalues found:  2" in out.getvalue() and "Success" in out.getvalue() and out.getvalue().count("residual") == 3


def test_feast_set_defaults_and_validation():
    import feastcuda as fc
    fpm = fc.feastinit()
    assert fc.feast_set_defaults(fpm, print_level=0, integration_points=16, tolerance_exp=10, max_refinement=5) is fpm
    assert fpm[:4] == [0, 16, 10, 5]
    for bad in (dict(print_level=2), dict(integration_points=0), dict(tolerance_exp=17), dict(max_refinement=0)):
        with pytest.raises(ValueError):
            fc.feast_set_defaults(fc.feastinit(), **bad)
    with pytest.raises(ValueError):
        fc.feast_set_defaults([0] * 10)


def test_rational_filter_values():
    """feast_rational / feast_rationalx (core/feast_tools.jl:483-540): ~1 inside the interval, ~0 outside, 1/2 at the end points;
    feast_grational(x) on the full contour likewise; both argument orders of the x forms."""
    import feastcuda as fc
    fpm = fc.feastinit()
    Emin, Emax = 1.0, 3.0
    lam = np.array([0.0, 1.0, 1.5, 2.0, 2.9, 3.0, 4.0, 10.0])
    f = fc.feast_rational(lam, Emin, Emax, fpm)
    Z, W = fo.feast_contour(Emin, Emax, fo.feastdefault(fo.feastinit()))
    want = np.array([2 * sum((w / (z - x)).real for z, w in zip(Z, W)) for x in lam])
    assert np.allclose(f, want, atol=1e-14)
    assert abs(f[3] - 1.0) < 1e-6 and abs(f[1] - 0.5) < 1e-3 and abs(f[5] - 0.5) < 1e-3 and abs(f[0]) < 0.05 and abs(f[7]) < 1e-8
    assert np.allclose(fc.feast_rationalx(Z, W, lam), f) and np.allclose(fc.feast_rationalx(lam, Z, W), f)
    assert fc.feast_rational_expert is fc.feast_rationalx
    lamc = np.array([0.2 + 0.1j, 0.9j, 3.0 + 0j])
    g = fc.feast_grational(lamc, 0j, 1.0, fc.feastinit())
    Zg, Wg = fo.feast_gcontour(0j, 1.0, fo.feastdefault(fo.feastinit()))
    assert np.allclose(g, [sum(w / (z - x) for z, w in zip(Zg, Wg)) for x in lamc], atol=1e-14)
    assert abs(g[0] - 1.0) < 1e-6 and abs(g[2]) < 1e-6
    assert np.allclose(fc.feast_grationalx(lamc, Zg, Wg), g)
    with pytest.raises(ValueError):
        fc.feast_rationalx(Z, W[:-1], lam)


def test_customcontour_weights_match_the_oracle():
    """feast_customcontour (core/feast_tools.jl:378-398): W_i = (Z_{i+1} - Z_{i-1}) / (2 ne), fpm[2] = ne."""
    import feastcuda as fc
    theta = 2 * np.pi * (np.arange(12) + 0.5) / 12
    nodes = 2.0 + 1.5 * np.exp(1j * theta)
    fpm = fc.feastinit()
    Z, W = fc.feast_customcontour(nodes, fpm)
    Zo, Wo = fo.feast_customcontour(nodes, fo.feastinit())
    assert fpm[1] == 12 and np.allclose(Z, Zo) and np.allclose(W, Wo)
    with pytest.raises(ValueError):
        fc.feast_customcontour([], fc.feastinit())


def test_backend_keywords_follow_the_reference_rules():
    """test/test_backend_api.jl:25-66: `parallel` and `backend` must agree, unknown names and unavailable explicit backends are
    ArgumentErrors (raised before any device is touched), :auto may fall back."""
    import feastcuda as fc
    from feastcuda.api import _normalize_backend, _select_backend
    assert _normalize_backend(None, None) == "serial" and _normalize_backend(True, None) == "auto" and _normalize_backend(False, None) == "serial"
    assert _normalize_backend("serial", "serial") == "serial" and _normalize_backend(None, ":mpi") == "mpi"
    with pytest.raises(ValueError, match="Conflicting"):
        _normalize_backend("threads", "serial")
    with pytest.raises(ValueError, match="Unknown backend"):
        _normalize_backend(None, "bogus")
    A = np.diag(2.0 * np.ones(10)) - np.diag(np.ones(9), 1) - np.diag(np.ones(9), -1)
    B = np.eye(10)
    for bad in (dict(backend="serial", parallel="threads"), dict(backend="bogus"), dict(backend="threads"), dict(parallel="threads"),
                dict(backend="mpi"), dict(backend="distributed")):
        with pytest.raises(ValueError):
            fc.feast(A, B, (0.1, 3.9), M0=10, fpm=fc.feastinit(), **bad)
    with pytest.raises(ValueError):
        fc.feast_general(A.astype(complex), 0j, 3.0, M0=10, fpm=fc.feastinit(), backend="threads")
    # fallbacks: :auto, parallel=true, and a non-strict legacy request resolve to the engine without raising
    for ok in (dict(backend="auto"), dict(parallel=True), dict(backend="serial"), dict(parallel=False)):
        kw = dict(ok, comm=None, use_threads=None)
        assert _select_backend(kw) in ("serial", "auto") and kw == {}
    with pytest.raises(ValueError):
        _select_backend(dict(backend="auto", strict_backend=True, parallel="threads"))


def test_tiled_host_passes_over_dense_operators():
    """feastcuda/_host.py: the tiled symmetry checks and the tiled column-major copy agree with the plain NumPy forms they replace
    (these two passes were 3.8 s of the 4.8 s end-to-end time of configs[1])."""
    import feastcuda as fc
    from feastcuda._host import column_major
    rng = np.random.default_rng(11)
    n = 1100                                   # > 2 tiles: the threaded path
    S = rng.standard_normal((n, n))
    S = S + S.T
    H = S + 1j * (lambda K: K - K.T)(rng.standard_normal((n, n)))
    assert fc.issymmetric(S) and fc.ishermitian(S) and fc.ishermitian(H) and not fc.issymmetric(H)
    assert fc.issymmetric(np.asfortranarray(S)) and fc.ishermitian(np.asfortranarray(H))
    for (i, j) in ((3, 1099), (1099, 3), (600, 601), (1050, 520)):      # one entry off, in different tiles
        T = S.copy()
        T[i, j] += 1e-13
        assert not fc.issymmetric(T) and not fc.ishermitian(T)
        G = H.copy()
        G[i, j] += 1e-13j
        assert not fc.ishermitian(G)
    assert not fc.issymmetric(rng.standard_normal((4, 5))) and fc.issymmetric(np.eye(3)) and not fc.ishermitian(np.eye(3) * 1j)
    R = rng.standard_normal((n, n + 7))
    F = column_major(R, np.float64)
    assert F.flags.f_contiguous and np.array_equal(F, R) and column_major(F, np.float64) is F
    Z = column_major(R[:40, :30], np.complex128)
    assert Z.flags.f_contiguous and Z.dtype == np.complex128 and np.array_equal(Z.real, R[:40, :30])
