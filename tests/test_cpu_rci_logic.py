"""CPU tests of the reverse-communication state machines' HOST logic (feastcuda/rci.py) with a NumPy stand-in for the engine's
stage-level entry points (accumulate / gram / rowtransform / eig_general are GPU-tested one by one in tests/test_gpu_stages.py
and driven through the real engine in tests/test_gpu_solve.py).  The stand-in lives here, in test code: it checks the job-code
sequencing, the Ref/fpm bookkeeping and the arithmetic order of kernel/feast_kernel.jl, not the device kernels."""
import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import feast_oracle as fo


class NumpyStages:
    """What the four stage calls compute, in NumPy (reference arithmetic: kernel/feast_kernel.jl:143-187, 766-845)."""

    def accumulate(self, w, Y, Q):
        return Q + w * Y

    def gram(self, X, Y):
        return X.conj().T @ Y

    def rowtransform(self, X, T):
        return X @ T

    def eig_general(self, S, B=None):
        lam, V = sla.eig(S) if B is None else sla.eig(S, B)
        V = V / np.linalg.norm(V, axis=0)
        return lam, V


class RankRevealingStages(NumpyStages):
    """eig_general as feastcuda_eig_general does it for FEAST moment pencils: eigenpairs on the numerical range of B, the null
    directions reported as +inf (LAPACK's QZ may return an arbitrary value for them)."""

    def eig_general(self, S, B=None):
        U, sv, Vh = np.linalg.svd(B)
        k = int((sv > 1e-13 * sv[0]).sum())
        lam = np.full(S.shape[0], np.inf, dtype=complex)
        V = np.zeros_like(S, dtype=complex)
        l, y = sla.eig(U[:, :k].conj().T @ S @ Vh[:k].conj().T, np.diag(sv[:k]).astype(complex))
        lam[:k] = l
        V[:, :k] = Vh[:k].conj().T @ y
        V[:, k:] = Vh[k:].conj().T
        return lam, V / np.linalg.norm(V, axis=0)


def _drive_general(A, B, Emid, r, M0, fpm=None, maxiter=2000):
    """The caller's side of feast_grci!: sparse LU per node (factorize), block solves, B q and A q products."""
    import feastcuda as fc
    N = A.shape[0]
    Bm = sp.identity(N, dtype=complex, format="csc") if B is None else sp.csc_matrix(B, dtype=complex)
    Am = sp.csc_matrix(A, dtype=complex)
    ijob, Ze, eps, loop, mode, info = fc.Ref(-1), fc.Ref(0j), fc.Ref(0.0), fc.Ref(0), fc.Ref(0), fc.Ref(-1)
    work, workc = np.zeros((N, M0)), np.zeros((N, M0), dtype=complex)
    Aq, Sq = np.zeros((M0, M0), dtype=complex), np.zeros((M0, M0), dtype=complex)
    lam, q, res = np.zeros(M0, dtype=complex), np.zeros((N, M0), dtype=complex), np.zeros(M0)
    fpm = fc.feastinit() if fpm is None else fpm
    state = fc.FeastRCIState()
    eng = NumpyStages()
    lu, jobs = None, []
    for _ in range(maxiter):
        fc.feast_grci(ijob, N, Ze, work, workc, Aq, Sq, fpm, eps, loop, Emid, r, M0, lam, q, mode, res, info, state=state, engine=eng)
        jobs.append(ijob.v)
        if ijob.v == 10:
            lu = spla.splu((Ze.v * Bm - Am).tocsc())
        elif ijob.v == 11:
            workc[:, :M0] = lu.solve(np.asarray(Bm @ workc[:, :M0]))
        elif ijob.v == 40:
            workc[:, :mode.v] = Bm @ q[:, :mode.v]
        elif ijob.v == 30:
            workc[:, :mode.v] = Am @ q[:, :mode.v]
        elif ijob.v == 0:
            break
    return lam[:mode.v].copy(), q[:, :mode.v].copy(), res[:mode.v].copy(), info.v, loop.v, eps.v, jobs, state


def test_grci_full_loop_matches_lapack_and_the_oracle():
    """feast_grci! (kernel/feast_kernel.jl:646-962) driven to convergence on a non-Hermitian pencil: job sequence
    10,11 (x ne), 40, 30, 30, then 0 or 10 again; eigenvalues inside the circle against LAPACK and the oracle's feast_general."""
    rng = np.random.default_rng(2)
    n, M0 = 60, 12
    d = np.linspace(-1.0, 3.0, n) + 0.3j * np.sin(np.arange(n))
    A = np.diag(d) + 0.02 * (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    B = np.diag(1.0 + 0.1 * rng.random(n)).astype(complex)
    Emid, r = 0.4 + 0.1j, 0.33
    fpm0 = fo.feastdefault(fo.feastinit())
    for Bm in (None, B):
        w = sla.eigvals(A) if Bm is None else sla.eigvals(A, Bm)
        want = [x for x in w if fo.feast_inside_gcontour(x, Emid, r, fpm0)]
        assert 3 <= len(want) <= M0 - 2
        lam, q, res, info, loops, eps, jobs, st = _drive_general(A, Bm, Emid, r, M0)
        assert info == 0 and len(lam) == len(want)
        for x in lam:
            assert min(abs(x - y) for y in want) < 1e-9
        assert np.all(np.diff(np.abs(lam)) >= -1e-15)                      # feast_sort_general!: ascending |lambda|
        Bq = q if Bm is None else Bm @ q
        for j in range(len(lam)):
            assert np.linalg.norm(A @ q[:, j] - lam[j] * Bq[:, j]) < 1e-9 * max(1.0, abs(lam[j]))
            assert abs(np.linalg.norm(q[:, j]) - 1.0) < 1e-12
        ne = 16
        first = jobs[:2 * ne + 3]
        assert first == [10, 11] * ne + [40, 30, 30]
        assert jobs[-1] == 0 and not st.initialized
        ro = fo.feast_general(A, Bm, Emid, r, M0, fo.feastinit())
        assert ro.M == len(lam)


def test_grci_argument_codes_and_empty_contour():
    import feastcuda as fc
    N, M0 = 8, 3
    mk = lambda: (np.zeros((N, M0)), np.zeros((N, M0), dtype=complex), np.zeros((M0, M0), dtype=complex), np.zeros((M0, M0), dtype=complex),
                  np.zeros(M0, dtype=complex), np.zeros((N, M0), dtype=complex), np.zeros(M0))
    for args, code in (((0j, 0.0), 4), ((0j, -1.0), 4)):
        work, workc, Aq, Sq, lam, q, res = mk()
        info = fc.Ref(-1)
        fc.feast_grci(fc.Ref(-1), N, fc.Ref(0j), work, workc, Aq, Sq, fc.feastinit(), fc.Ref(0.0), fc.Ref(0), args[0], args[1], M0, lam, q,
                      fc.Ref(0), res, info)
        assert info.v == code                                              # Feast_ERROR_EMID_R
    # a contour that holds no eigenvalue ends with info = 5 (Feast_ERROR_NO_CONVERGENCE) after the first projection
    A = np.diag(np.arange(1.0, N + 1)).astype(complex)
    lam, q, res, info, loops, eps, jobs, st = _drive_general(A, None, 20.0 + 0j, 0.5, M0)
    assert info == 5 and st.M == 0 and jobs[-1] == 0 and not st.initialized


def test_srci_logic_with_the_stand_in_engine():
    """feast_srci! host logic on CPU: the same drive as tests/test_gpu_solve.py::test_rci_state_machines_drive_a_full_solve with the
    NumPy stand-in -- rank-deficient moment matrices (M0 = 8, four eigenvalues inside) included."""
    import feastcuda as fc
    L1 = fo.laplacian_1d(10).tocsc()
    N, M0, Emin, Emax = 10, 8, 0.1, 1.9


    ijob, Ze, eps, loop, mode, info = fc.Ref(-1), fc.Ref(0j), fc.Ref(0.0), fc.Ref(0), fc.Ref(0), fc.Ref(-1)
    work, workc = np.zeros((N, M0)), np.zeros((N, M0), dtype=complex)
    Aq, Sq = np.zeros((M0, M0)), np.zeros((M0, M0))
    lam, q, res = np.zeros(M0), np.zeros((N, M0)), np.zeros(M0)
    fpm, state, eng, z = fc.feastinit(), fc.FeastRCIState(), RankRevealingStages(), 0j
    Ad = L1.toarray().astype(complex)
    for _ in range(400):
        fc.feast_srci(ijob, N, Ze, work, workc, Aq, Sq, fpm, eps, loop, Emin, Emax, M0, lam, q, mode, res, info, state=state, engine=eng)
        if ijob.v == 10:
            z = Ze.v
        elif ijob.v == 11:
            workc[:, :M0] = np.linalg.solve(z * np.eye(N) - Ad, work[:, :M0].astype(complex))
        elif ijob.v == 30:
            work[:, :mode.v] = (Ad @ q[:, :mode.v]).real
        elif ijob.v == 0:
            break
    w = np.linalg.eigvalsh(L1.toarray())
    want = w[(w >= Emin) & (w <= Emax)]
    assert info.v == 0 and mode.v == len(want) and np.allclose(lam[:mode.v], want, atol=1e-9) and res[:mode.v].max() < 1e-10


def test_custom_contour_rci_names_use_the_callers_nodes():
    """feast_srcix!/feast_grcix! (kernel/feast_kernel.jl:294-336): the caller's (Zne, Wne) replace the default quadrature -- a
    16-node Gauss contour passed explicitly gives the state machine 16 nodes and the same eigenvalues; ifeast_* are the same
    functions."""
    import feastcuda as fc
    assert fc.ifeast_srci is fc.feast_srci and fc.ifeast_hrci is fc.feast_hrci and fc.ifeast_grci is fc.feast_grci
    N, M0, Emin, Emax = 10, 6, 0.1, 1.9
    L1 = fo.laplacian_1d(N).toarray().astype(complex)
    fpm16 = fo.feastinit()
    fpm16[1] = 16
    Z, W = fo.feast_contour(Emin, Emax, fo.feastdefault(fpm16))
    ijob, Ze, eps, loop, mode, info = fc.Ref(-1), fc.Ref(0j), fc.Ref(0.0), fc.Ref(0), fc.Ref(0), fc.Ref(-1)
    work, workc = np.zeros((N, M0)), np.zeros((N, M0), dtype=complex)
    Aq, Sq = np.zeros((M0, M0)), np.zeros((M0, M0))
    lam, q, res = np.zeros(M0), np.zeros((N, M0)), np.zeros(M0)
    fpm, state, eng, z, nodes = fc.feastinit(), fc.FeastRCIState(), RankRevealingStages(), 0j, []
    for _ in range(400):
        fc.feast_srcix(ijob, N, Ze, work, workc, Aq, Sq, fpm, eps, loop, Emin, Emax, M0, lam, q, mode, res, info, Z, W, state=state, engine=eng)
        if ijob.v == 10:
            z = Ze.v
            nodes.append(z)
        elif ijob.v == 11:
            workc[:, :M0] = np.linalg.solve(z * np.eye(N) - L1, work[:, :M0].astype(complex))
        elif ijob.v == 30:
            work[:, :mode.v] = (L1 @ q[:, :mode.v]).real
        elif ijob.v == 0:
            break
    w = np.linalg.eigvalsh(L1.real)
    want = w[(w >= Emin) & (w <= Emax)]
    assert info.v == 0 and state.ne == 16 and np.allclose(nodes[:16], Z) and np.allclose(lam[:mode.v], want, atol=1e-9)
    import pytest
    with pytest.raises(ValueError):
        fc.feast_srcix(fc.Ref(-1), N, fc.Ref(0j), work, workc, Aq, Sq, fc.feastinit(), fc.Ref(0.0), fc.Ref(0), Emin, Emax, M0, lam, q,
                       fc.Ref(0), res, fc.Ref(0), Z, W[:-1])


def test_hrci_first_sweep_reproduces_the_reference_moments():
    """feast_hrci! (kernel/feast_kernel.jl:397-644) accumulates 2 w_e Y without a Hermitian part: after one contour sweep the
    moments are Q0^H g(A) Q0 and Q0^H h(A) Q0 with g = sum 2w/(z-x), h = sum 2wz/(z-x) -- restated here and compared."""
    import feastcuda as fc
    v = np.array([0.1 + 0.2j, -0.05 + 0.1j])
    Ah = (np.diag([2.0, 3.0, 4.0]).astype(complex) + np.diag(v, 1) + np.diag(np.conj(v), -1))
    N = M0 = 3
    Emin, Emax = 1.5, 4.5
    ijob, Ze, eps, loop, mode, info = fc.Ref(-1), fc.Ref(0j), fc.Ref(0.0), fc.Ref(0), fc.Ref(0), fc.Ref(-1)
    work, workc = np.zeros((N, M0)), np.zeros((N, M0), dtype=complex)
    zAq, zSq = np.zeros((M0, M0), dtype=complex), np.zeros((M0, M0), dtype=complex)
    lam, q, res = np.zeros(M0), np.zeros((N, M0), dtype=complex), np.zeros(M0)
    fpm, state, eng, z = fc.feastinit(), fc.FeastRCIState(), NumpyStages(), 0j
    for _ in range(100):
        fc.feast_hrci(ijob, N, Ze, work, workc, zAq, zSq, fpm, eps, loop, Emin, Emax, M0, lam, q, mode, res, info, state=state, engine=eng)
        if ijob.v == 10:
            z = Ze.v
        elif ijob.v == 11:
            workc[:, :M0] = np.linalg.solve(z * np.eye(N) - Ah, workc[:, :M0])
        elif ijob.v in (30, 0):
            break
    Z, W = fo.feast_contour(Emin, Emax, fo.feastdefault(fo.feastinit()))
    G = sum(2 * w * np.linalg.inv(zz * np.eye(N) - Ah) for zz, w in zip(Z, W))
    H = sum(2 * w * zz * np.linalg.inv(zz * np.eye(N) - Ah) for zz, w in zip(Z, W))
    Q0 = state.Q0
    assert ijob.v == 30 and np.allclose(zAq, Q0.conj().T @ G @ Q0, atol=1e-12) and np.allclose(zSq, Q0.conj().T @ H @ Q0, atol=1e-12)
    want = np.sort(sla.eigvals(Q0.conj().T @ H @ Q0, Q0.conj().T @ G @ Q0).real)
    assert np.allclose(np.sort(lam[:mode.v]), want[(want >= Emin) & (want <= Emax)], atol=1e-10)
    # the oracle's restatement of the whole routine (feast_hmom) gives the same first sweep from the same Q0
    ro, oAq, oSq = fo.feast_hmom(Ah, None, Emin, Emax, M0, fo.feastinit(), Q0=Q0, max_sweeps=1)
    assert np.allclose(oAq, zAq, atol=1e-12) and np.allclose(oSq, zSq, atol=1e-12) and ro.M == mode.v
    assert np.allclose(np.sort(ro.lambda_), np.sort(lam[:mode.v]), atol=1e-10)
    # ... and documents the defect: only the eigenvalue at the centre of the interval (3.0 here) comes out right
    exact = np.linalg.eigvalsh(Ah)
    errs = np.abs(np.sort(ro.lambda_) - exact)
    assert errs[1] < 0.05 and errs[0] > 0.1 and errs[2] > 0.1


class BandStandIn(RankRevealingStages):
    """NumPy stand-in for the engine calls the S-MOM banded route makes (set_band / clear_b / apply / block_solve / stats)."""

    def __init__(self):
        self.ops = {}

    def set_band(self, which, AB, k, structure):
        self.ops[which] = fo.banded_to_full(np.asarray(AB, dtype=complex), k, hermitian=True)

    def clear_b(self):
        self.ops.pop(1, None)

    def apply(self, which, X):
        return self.ops[which] @ X

    def block_solve(self, z, rhs, **kw):
        n = rhs.shape[0]
        Bm = self.ops.get(1, np.eye(n))
        return np.linalg.solve(z * Bm - self.ops[0], rhs), None, None

    def stats(self):
        return {}


def test_banded_smom_route_follows_the_reference_driver():
    """feast_sbev!/feast_sbgv! the reference's way (banded/feast_banded.jl:9-186: feast_srci! + band LU per node), opt-in in the
    mirror (method="smom"): same eigenpairs and loop count as the oracle's restatement of that driver (oracle feast_sbev -> S-MOM)."""
    import feastcuda as fc
    rng = np.random.default_rng(4)
    n, k, M0 = 60, 2, 10
    A = np.diag(np.linspace(1.0, 12.0, n))
    for d in range(1, k + 1):
        v = 0.1 * rng.standard_normal(n - d)
        A += np.diag(v, d) + np.diag(v, -d)
    w = np.linalg.eigvalsh(A)
    Emin, Emax = 0.5 * (w[9] + w[10]), 0.5 * (w[15] + w[16])
    AB = fo.full_to_banded(A, k)
    Q0 = fo.seeded_subspace(n, M0, complex_storage=False)
    r = fc.feast_sbev(AB, k, Emin, Emax, M0, fc.feastinit(), method="smom", engine=BandStandIn(), Q0=Q0)
    ro = fo.feast_smom(A, None, Emin, Emax, M0, fo.feastinit(), Q0=Q0)
    assert r.info == 0 and r.M == 6
    assert np.allclose(r.lambda_, w[10:16], atol=1e-10) and r.res.max() < 1e-10
    # M0 = 10 > 6 eigenvalues inside makes the moment pencil rank deficient: LAPACK's QZ (the reference, the oracle) then returns
    # arbitrary values for the null directions, some inside the interval, and loses digits on the true ones (here the oracle
    # reports M = 8 and 3.8591 for the eigenvalue 3.8617); the engine's rank-revealing reduced solve does not.
    assert ro.info == 0 and ro.M >= r.M
    # with M0 equal to the eigenvalue count the pencil is regular and both routes agree value for value and loop for loop
    Q6 = fo.seeded_subspace(n, 6, complex_storage=False)
    r6 = fc.feast_sbev(AB, k, Emin, Emax, 6, fc.feastinit(), method="smom", engine=BandStandIn(), Q0=Q6)
    o6 = fo.feast_smom(A, None, Emin, Emax, 6, fo.feastinit(), Q0=Q6)
    assert r6.info == o6.info == 0 and r6.M == o6.M == 6 and r6.loop == o6.loop
    assert np.allclose(r6.lambda_, np.sort(o6.lambda_), atol=1e-10) and np.allclose(r6.lambda_, w[10:16], atol=1e-10)
    assert fo.subspace_angle(r6.q.astype(complex), np.asarray(o6.q, dtype=complex)) < 1e-8
    for j in range(r.M):
        assert np.linalg.norm(A @ r.q[:, j] - r.lambda_[j] * r.q[:, j]) < 1e-9 * np.linalg.norm(r.q[:, j])
    assert r.q.dtype == np.float64 and np.all(np.diff(r.lambda_) > 0)
    # argument errors are FeastError codes of the RCI kernel or host-side exceptions, as in the reference
    import pytest
    with pytest.raises(ValueError):
        fc.feast_sbev(AB, k, Emax, Emin, M0, fc.feastinit(), method="smom", engine=BandStandIn())
