"""CPU: pins the oracle (oracle/feast_oracle.py, oracle/feast_port.py) against the known-answer cases of the reference's
own test-suite (tests/golden/reference_known_answers.json) and the committed regression anchors."""
import json
from pathlib import Path

import numpy as np
import pytest
import scipy.linalg as sla
import scipy.sparse as sp

import feast_oracle as fo
import feast_port as fp

G = Path(__file__).resolve().parent / "golden"
KA = json.loads((G / "reference_known_answers.json").read_text())
REG = json.loads((G / "engine_regression.json").read_text())


def _c(d):
    return np.array(d["re"]) + 1j * np.array(d["im"])


def test_feastinit_sentinel_and_defaults():
    """runtests.jl:10-70: -111 sentinel; defaults fpm[2]=8, fpm[3]=12, fpm[4]=20, fpm[8]=16; contour lengths."""
    fpm = fo.feastinit()
    assert len(fpm) == 64 and all(v == -111 for v in fpm)
    fo.feastdefault(fpm)
    assert (fpm[1], fpm[2], fpm[3], fpm[7], fpm[15], fpm[17]) == (8, 12, 20, 16, 0, 100)
    assert fpm == REG["feastdefault"]
    Z, W = fo.feast_contour(0.0, 1.0, fo.feastinit())
    assert len(Z) == len(W) == 8
    Zg, Wg = fo.feast_gcontour(0.0, 1.0, fo.feastinit())
    assert len(Zg) == len(Wg) == 16
    with pytest.raises(ValueError):
        fo.check_feast_srci_input(10, 0, 0.0, 1.0, fpm)
    with pytest.raises(ValueError):
        fo.check_feast_srci_input(10, 4, 1.0, 0.0, fpm)


def test_contour_golden_numbers():
    k = KA["contour_default"]
    Z, W = fo.feast_contour(k["Emin"], k["Emax"], fo.feastinit())
    assert abs(Z[0] - complex(*k["Z1"])) < 1e-15 and abs(W[0] - complex(*k["W1"])) < 1e-16
    assert np.allclose(Z, _c(REG["contour_0.5_1.5"]["Z"]), rtol=0, atol=1e-15)
    assert np.allclose(W, _c(REG["contour_0.5_1.5"]["W"]), rtol=0, atol=1e-15)
    # the half-contour rule integrates the indicator: sum 2 Re(w/(z - x)) = 1 inside, ~0 far outside
    rho = lambda x: np.real(np.sum(2 * W / (Z - x)))
    assert abs(rho(1.0) - 1.0) < 1e-6 and abs(rho(3.0)) < 1e-6
    Zg, Wg = fo.feast_gcontour(0.0, 2.0, fo.feastinit())
    assert np.allclose(Zg, _c(REG["gcontour_0_2"]["Z"]), atol=1e-15) and np.allclose(Wg, _c(REG["gcontour_0_2"]["W"]), atol=1e-15)
    g = lambda x: np.sum(Wg / (Zg - x))
    assert abs(g(0.3 + 0.2j) - 1.0) < 1e-6 and abs(g(5.0)) < 1e-6


def test_helper_known_answers():
    k = KA["KA12_reorder"]
    lam = np.array(k["lambda"])
    vec = np.array(k["vectors"], dtype=complex)
    src = vec.copy()
    m = fo.reorder_by_interval(lam, vec, *k["interval"], len(lam))
    assert m == k["m"] and lam.tolist() == k["lambda_out"] and np.array_equal(vec, src[:, k["perm"]])
    k = KA["KA12_sort"]
    lam, q, res = np.array(k["lambda"]), np.array(k["q"]), np.array(k["res"])
    q0 = q.copy()
    fo.feast_sort(lam, q, res, 4)
    assert lam.tolist() == [1.0, 2.0, 3.0, 4.0] and np.array_equal(q, q0[:, k["perm"]]) and res.tolist() == [0.1, 0.2, 0.3, 0.4]
    k = KA["KA12_sort_general"]
    lam = _c(k["lambda"])
    lam0 = lam.copy()
    q = np.array(KA["KA12_sort"]["q"], dtype=complex)
    res = np.array(k["res"])
    fo.feast_sort_general(lam, q, res, 4)
    assert np.array_equal(lam, lam0[k["perm"]]) and res.tolist() == [0.1, 0.2, 0.3, 0.4]
    k = KA["KA12_residual"]
    A, B, q, lam = np.array(k["A"]), np.diag(k["B_diag"]), np.array(k["q"]), np.array(k["lambda"])
    want = [np.linalg.norm(A @ q[:, j] - lam[j] * (B @ q[:, j])) / max(abs(lam[j]), 1.0) for j in range(2)]
    assert np.allclose(fo.feast_residual(A, B, lam, q, 2), want)
    k = KA["KA12_qr_compress"]
    src = _c(k["src"])
    Q, rank = fo.qr_compress(src, 4, rank_tol=np.sqrt(np.finfo(float).eps))
    assert rank == k["rank"]
    assert np.allclose(Q.conj().T @ Q, np.eye(rank), atol=1e-12) and np.linalg.norm(src - Q @ (Q.conj().T @ src)) <= 1e-12


def test_ka1_ka2_ka3_dense_and_sparse_hermitian():
    k = KA["KA1"]
    A = np.array(k["A"])
    r = fo.feast_sygv(A, np.eye(3), *k["interval"], k["M0"], fo.feastinit())
    assert r.info == 0 and r.M == 3 and np.allclose(np.sort(r.lambda_), k["expected"], atol=k["atol"])
    k = KA["KA2"]
    r = fo.feast_heev(_c(k["A"]), *k["interval"], k["M0"], fo.feastinit())
    assert r.info == 0 and r.M == 3 and np.allclose(np.sort(r.lambda_), k["expected"], atol=k["atol"])
    k = KA["KA3"]
    r = fo.feast_hcsrev(sp.csc_matrix(_c(k["A"])), *k["interval"], k["M0"], fo.feastinit())
    assert r.info == 0 and r.M == 3 and np.allclose(np.sort(r.lambda_), k["expected"], atol=k["atol"])


def test_ka4_general_standard_and_generalized():
    k = KA["KA4"]
    A, B = _c(k["A"]), _c(k["B"])
    r = fo.feast_general(A, None, complex(*k["center"]), k["radius"], k["M0"], fo.feastinit())
    assert r.info == 0 and r.M == 2 and np.allclose(np.sort(r.lambda_.real), k["expected_standard"], atol=k["atol"])
    r = fo.feast_general(sp.csc_matrix(A), None, complex(*k["center"]), k["radius"], k["M0"], fo.feastinit())
    assert r.info == 0 and r.M == 2 and np.allclose(np.sort(r.lambda_.real), k["expected_standard"], atol=k["atol"])
    rg = fo.feast_general(A, B, complex(*k["center"]), k["radius"], k["M0"], fo.feastinit(), residual="true")
    assert rg.M == 2 and np.allclose(np.sort(rg.lambda_.real), k["expected_generalized"], atol=k["atol"])


def test_ka5_ka6_ka10_sparse():
    k = KA["KA5"]
    A = fo.laplacian_1d(k["n"]).tocsc()
    r = fo.feast_scsrev(A, *k["interval"], k["M0"], fo.feastinit())
    assert r.info == 0 and r.M == 10 and np.allclose(np.sort(r.lambda_), k["expected"], atol=k["atol"])
    k = KA["KA6"]
    A, B = sp.diags(np.array(k["A_diag"], dtype=complex)).tocsc(), sp.diags(np.array(k["B_diag"], dtype=complex)).tocsc()
    r = fo.feast_hcsrgv(A, B, *k["interval"], k["M0"], fo.feastinit())
    assert r.info == 0 and r.M == len(k["expected"]) and np.allclose(np.sort(r.lambda_), k["expected"], atol=k["atol"])
    k = KA["KA10"]
    r = fo.feast_scsrev(sp.diags(k["A_diag"]).tocsc(), *k["interval"], k["M0"], fo.feastinit())
    assert r.info == 0 and r.M == 3 and np.allclose(np.sort(r.lambda_), k["expected"], atol=k["atol"])
    rd = fo.feast_syev(np.diag(k["A_diag"]), *k["interval"], k["M0"], fo.feastinit())
    assert rd.M == 3 and np.allclose(np.sort(rd.lambda_), k["expected"], atol=k["atol"])


def test_ka11_rank_compression_dense_sparse_banded():
    k = KA["KA11"]
    n = k["n"]
    d = np.arange(1.0, n + 1)

    def fpm():
        f = fo.feastinit()
        for kk, v in k["fpm"].items():
            f[int(kk) - 1] = v
        return f
    for r in (fo.feast_syev(np.diag(d), *k["interval"], k["M0"], fpm()),
              fo.feast_scsrev(sp.diags(d).tocsc(), *k["interval"], k["M0"], fpm()),
              fo.feast_hbev(d.reshape(1, n).astype(complex), 0, *k["interval"], k["M0"], fpm())):
        assert r.info == 0 and r.M == 2
        assert np.allclose(np.sort(r.lambda_), k["expected"], atol=k["atol"]) and r.res.max() < k["max_res"]


def test_ka8_banded_real_symmetric_moment_solver():
    """runtests.jl:605-636: 1-D Laplacian n=8 in band storage, (0.5, 3.1) -- S-MOM driver."""
    n = 8
    A = fo.laplacian_1d(n).toarray()
    AB = fo.full_to_banded(A, 1)
    assert np.allclose(fo.banded_to_full(AB, 1, hermitian=False), A)   # converter round trip, runtests.jl:582-600
    r = fo.feast_sbev(AB, 1, 0.5, 3.1, 8, fo.feastinit())
    w = np.linalg.eigvalsh(A)
    want = w[(w >= 0.5) & (w <= 3.1)]
    assert r.info == 0 and r.M == len(want) and np.allclose(np.sort(r.lambda_), want, atol=1e-8)


def test_ka7_direct_equals_gmres():
    """runtests.jl:306-395,442-508: the iterative path agrees with the direct one."""
    A = fo.laplacian_1d(12).tocsc()
    d = fo.feast_scsrgv(A, sp.identity(12, format="csc"), 0.1, 1.2, 6, fo.feastinit())
    g = fo.feast_scsrgv(A, sp.identity(12, format="csc"), 0.1, 1.2, 6, fo.feastinit(), solver="gmres", solver_tol=1e-12,
                        solver_maxiter=400, solver_restart=30)
    assert d.info == g.info == 0 and d.M == g.M and np.allclose(np.sort(d.lambda_), np.sort(g.lambda_), atol=1e-8)


def test_regression_anchor_laplacian3d():
    for filt in ("true", "reference"):
        k = REG[f"laplacian3d_N12_{filt}"]
        A = fo.laplacian_3d(k["N"]).astype(float).tocsc()
        fpm = fo.feastinit()
        if filt == "reference":
            fpm[3] = 60
        r = fo.feast_scsrev(A, *k["interval"], k["M0"], fpm, Q0=fo.seeded_subspace(k["N"] ** 3, k["M0"]), filter=filt)
        assert (r.M, r.info, r.loop) == (k["M"], k["info"], k["loop"])
        assert np.allclose(np.sort(r.lambda_), k["lambda"], rtol=1e-12, atol=1e-13)
        assert np.allclose(np.sort(r.lambda_), k["analytic"], rtol=1e-10, atol=1e-12)


def test_engine_ports_reach_the_reference_eigenpairs():
    """The NumPy ports of the ENGINE's inner solvers (block BiCGStab, multi-shift Lanczos) converge to the oracle's pairs."""
    N, M0 = 10, 16
    A = fo.laplacian_3d(N).astype(float).tocsr()
    ev = fo.laplacian_3d_eigs(N)
    Emin, Emax = 0.0, 0.5 * (ev[6] + ev[7])
    Q0 = fo.seeded_subspace(N ** 3, M0, complex_storage=False)
    ro = fo.feast_scsrev(A.tocsc(), Emin, Emax, M0, fo.feastinit(), Q0=Q0.astype(complex), filter="true")
    rl = fp.feast_hrr_mslanczos(A, Emin, Emax, M0, fo.feastinit(), Q0, inner_rel=1e-3, inner_maxiter=1000)
    rb = fp.feast_hrr_bicgstab(A, None, Emin, Emax, M0, fo.feastinit(), Q0, inner_rel=1e-3, inner_maxiter=400)
    for r in (rl, rb):
        assert r.info == 0 and r.M == ro.M == 7
        assert np.abs(np.sort(r.lambda_) - np.sort(ro.lambda_)).max() < 1e-10
        assert r.res.max() < 1e-12
        assert fo.subspace_angle(np.asarray(r.q, dtype=complex), np.asarray(ro.q, dtype=complex)) < 1e-8
    # tight multi-shift solves reproduce the exact-solve loop count
    rt = fp.feast_hrr_mslanczos(A, Emin, Emax, M0, fo.feastinit(), Q0, inner_rel=1e-10, inner_maxiter=3000)
    assert rt.loop == ro.loop
    # fpm[42] "single-precision solver": FP32 Lanczos vectors inside the FP64 refinement loop reach the same pairs
    rm = fp.feast_hrr_mslanczos(A, Emin, Emax, M0, fo.feastinit(), Q0, inner_rel=1e-3, inner_maxiter=1000, adaptive=True, mixed=True)
    assert rm.info == 0 and rm.M == ro.M and rm.stats["fp32_sweeps"] >= 2
    assert np.abs(np.sort(rm.lambda_) - np.sort(ro.lambda_)).max() < 1e-10 and rm.res.max() < 1e-12
    assert fo.subspace_angle(np.asarray(rm.q, dtype=complex), np.asarray(ro.q, dtype=complex)) < 1e-8


def test_multishift_lanczos_filter_equals_direct_solves():
    """One filter application: V_k c from the Lanczos tridiagonal == sum_e Re(2 w_e (z_e I - A)^-1 q) by sparse LU."""
    import scipy.sparse.linalg as spla
    n, m = 300, 5
    rng = np.random.default_rng(1)
    A = sp.diags([rng.uniform(-1, 0, n - 1), rng.uniform(1, 9, n), np.zeros(n - 1)], [-1, 0, 1]).tocsr()
    A = ((A + A.T) * 0.5).tocsr()
    Z, W = fo.feast_contour(2.0, 3.0, fo.feastinit())
    Q = rng.standard_normal((n, m))
    want = np.zeros((n, m))
    for z, w in zip(Z, W):
        lu = spla.splu((z * sp.identity(n, dtype=complex, format="csc") - A.astype(complex)).tocsc())
        want += np.real(2 * w * lu.solve(Q.astype(complex)))
    got = fp.mslanczos_filter(A, Q, None, Z, W, 1e-13, 600)
    assert np.abs(got - want).max() < 1e-9 * np.abs(want).max()
    theta = np.linspace(2.1, 2.9, m)
    got2 = fp.mslanczos_filter(A, Q, theta, Z, W, 1e-13, 600)   # Ritz-guess form is the same operator
    assert np.abs(got2 - want).max() < 1e-9 * np.abs(want).max()


def test_mixed_precision_port_on_an_ill_conditioned_interval():
    """FP32 Lanczos vectors (fpm[42]) on a 1-D Laplacian whose interval sits at the bottom of the spectrum (shifted systems with
    condition ~1e5): the FP64 refinement loop still reaches the FP64 pairs, at the price of more Lanczos steps (loss of
    orthogonality delays the FP32 recurrence) -- the reason the engine keeps mixed precision opt-in."""
    n, M0 = 400, 12
    A = fo.laplacian_1d(n).tocsr().astype(float)
    w = 2 - 2 * np.cos(np.arange(1, n + 1) * np.pi / (n + 1))
    Emin, Emax = 0.0, 0.5 * (w[5] + w[6])
    Q0 = fo.seeded_subspace(n, M0, complex_storage=False)
    r64 = fp.feast_hrr_mslanczos(A, Emin, Emax, M0, fo.feastinit(), Q0, inner_rel=1e-3, inner_maxiter=2400, adaptive=True)
    r32 = fp.feast_hrr_mslanczos(A, Emin, Emax, M0, fo.feastinit(), Q0, inner_rel=1e-3, inner_maxiter=2400, adaptive=True, mixed=True)
    for r in (r64, r32):
        assert r.info == 0 and r.M == 6 and r.res.max() < 1e-12
        assert np.abs(np.sort(r.lambda_) - w[:6]).max() < 1e-12
    assert r32.stats["fp32_sweeps"] >= 1 and sum(r32.stats["lz_steps"]) >= sum(r64.stats["lz_steps"])


def test_ka15_complex_symmetric_pencil():
    """runtests.jl:241-268: the complex-symmetric wrappers must find eigvals(A, B) inside the circle (atol 1e-7).  The oracle's
    general solver (one-sided Rayleigh-Ritz) is what the engine routes these names through."""
    k = KA["KA15_complex_symmetric"]
    un = lambda c: np.array(c["re"]) + 1j * np.array(c["im"])
    A, B = un(k["A"]), np.diag(np.array(k["B_diag"], dtype=complex))
    assert np.array_equal(A, A.T) and not np.allclose(A, A.conj().T)          # symmetric, not Hermitian
    center = complex(*k["center"])
    for Bm, want in ((B, un(k["expected_generalized"])), (None, un(k["expected_standard"]))):
        r = fo.feast_general(A, Bm, center, k["radius"], k["M0"], fo.feastinit())
        assert r.info == 0 and r.M == len(want)
        for lam in r.lambda_:
            assert np.abs(want - lam).min() < k["atol"]


def test_generalized_multishift_lanczos_port_on_the_reduced_config3_pair():
    """The design for BASELINE configs[3] (complex Hermitian FEM stiffness/mass pair): one Lanczos recurrence in the B-inner
    product serves all 16 nodes; B^-1 by Jacobi-PCG.  Same pairs as the oracle's zfeast_hcsrgv! restatement."""
    import scipy.sparse as sp

    def k1(n, h):
        return sp.diags([-np.ones(n - 1), 2 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1]) / h

    def m1(n, h):
        return h * sp.diags([np.ones(n - 1), 4 * np.ones(n), np.ones(n - 1)], [-1, 0, 1]) / 6
    dims = (7, 6, 5)
    hs = [1.0 / (n + 1) for n in dims]
    Ks, Ms = [k1(n, h) for n, h in zip(dims, hs)], [m1(n, h) for n, h in zip(dims, hs)]
    K = sp.kron(sp.kron(Ks[0], Ms[1]), Ms[2]) + sp.kron(sp.kron(Ms[0], Ks[1]), Ms[2]) + sp.kron(sp.kron(Ms[0], Ms[1]), Ks[2])
    Mass = sp.kron(sp.kron(Ms[0], Ms[1]), Ms[2])
    n = K.shape[0]
    D = sp.diags(np.exp(1j * np.random.default_rng(7).uniform(0, 2 * np.pi, n)))
    A = (D @ K @ D.conj()).tocsr()
    B = (D @ Mass @ D.conj()).tocsr()
    A, B = ((A + A.conj().T) * 0.5).tocsr(), ((B + B.conj().T) * 0.5).tocsr()
    w = np.sort(sla.eigh(A.toarray(), B.toarray(), eigvals_only=True))
    want, M0 = 6, 16
    assert w[want] - w[want - 1] > 1e-6 * w[want]
    Emin, Emax = 0.0, 0.5 * (w[want - 1] + w[want])
    Q0 = fo.seeded_subspace(n, M0)
    fpm = fo.feastinit()
    fpm[1] = 16
    r = fp.feast_hrr_mslanczos_gen(A, B, Emin, Emax, M0, fpm, Q0)
    ro = fo.feast_hcsrgv(A.tocsc(), B.tocsc(), Emin, Emax, M0, fo.feastinit(), Q0=Q0)
    assert r.info == ro.info == 0 and r.M == ro.M == want
    assert np.abs(np.sort(r.lambda_) - w[:want]).max() < 1e-10 * w[want] and r.res.max() < 1e-12
    assert fo.subspace_angle(np.asarray(r.q), np.asarray(ro.q, dtype=complex)) < 1e-8
    assert r.loop <= 2 and max(r.stats["pcg_iters"]) < 80


def test_multishift_arnoldi_port_on_the_reduced_config4_pencil():
    """The design for BASELINE configs[4] (general complex pencil, circular contour, 24 nodes): one stored Arnoldi basis per
    column serves every node through small shifted Hessenberg solves (multi-shift FOM).  Same eigenvalues as the analytic
    spectrum of the Toeplitz-Kronecker pencil and as the oracle's general solver, true residuals below 1e-12."""
    dims = (8, 8, 6)
    coef = [(0.4 + 0.1j, 1.0 + 0.05j, 0.9 - 0.05j), (0.3 - 0.1j, 0.8 + 0.1j, 0.75 + 0.05j), (0.5 + 0.2j, 0.6 - 0.05j, 0.65 + 0.02j)]
    eps = 0.05
    toe = lambda n, a, b, c: sp.diags([b * np.ones(n - 1), a * np.ones(n), c * np.ones(n - 1)], [-1, 0, 1])
    I = [sp.identity(n) for n in dims]
    T = [toe(n, *abc) for n, abc in zip(dims, coef)]
    S = [toe(n, 0.0, abc[1], abc[2]) for n, abc in zip(dims, coef)]
    ksum = lambda X: sp.kron(sp.kron(X[0], I[1]), I[2]) + sp.kron(sp.kron(I[0], X[1]), I[2]) + sp.kron(sp.kron(I[0], I[1]), X[2])
    A = ksum(T).tocsr()
    B = (sp.identity(A.shape[0]) + eps * ksum(S)).tocsr()
    n = A.shape[0]
    la, ls = [], []
    for nn, (a, b, c) in zip(dims, coef):
        th = np.arange(1, nn + 1) * np.pi / (nn + 1)
        la.append(a + 2 * np.sqrt(b * c) * np.cos(th))
        ls.append(2 * np.sqrt(b * c) * np.cos(th))
    lamA = (la[0][:, None, None] + la[1][None, :, None] + la[2][None, None, :]).ravel()
    lamS = (ls[0][:, None, None] + ls[1][None, :, None] + ls[2][None, None, :]).ravel()
    lam = lamA / (1 + eps * lamS)
    order = np.argsort(lam.real)
    Emid = complex(lam[order[0]].real, lam[order[:17]].imag.mean())
    dist = np.abs(lam - Emid)
    rad = 0.5 * (np.sort(dist)[11] + np.sort(dist)[12])
    inside = lam[dist <= rad]
    assert len(inside) == 12 and np.abs(dist - rad).min() > 5e-3
    M0 = 24
    Q0 = fo.seeded_subspace(n, M0)
    fpm = fo.feastinit()
    fpm[7] = 24
    r = fp.feast_general_msarnoldi(A, B, Emid, rad, M0, fpm, Q0)
    assert r.info == 0 and r.M == 12 and r.res.max() < 1e-12 and r.loop <= 3
    assert max(min(abs(g - x) for x in inside) for g in r.lambda_) < 1e-10
    assert np.all(np.diff(np.abs(r.lambda_)) >= -1e-14)                      # feast_sort_general!: ascending |lambda|
    Ad, Bd = A.toarray(), B.toarray()
    for j in range(r.M):
        x = r.q[:, j]
        assert np.linalg.norm(Ad @ x - r.lambda_[j] * (Bd @ x)) < 1e-11 * max(1.0, abs(r.lambda_[j]))
    ro = fo.feast_general(A.tocsc(), B.tocsc(), Emid, rad, M0, list(fpm), Q0=Q0, residual="true")
    assert ro.M == r.M


def test_chebyshev_generalized_lanczos_port_matches_the_oracle():
    """The generalized filter AS THE ENGINE RUNS IT (fixed Chebyshev polynomial for B^-1, unnormalised vectors, P^-1 inner product):
    same pairs as the oracle's zfeast_hcsrgv! restatement on a small FEM pair."""
    import scipy.sparse as sp

    def k1(n, h):
        return sp.diags([-np.ones(n - 1), 2 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1]) / h

    def m1(n, h):
        return h * sp.diags([np.ones(n - 1), 4 * np.ones(n), np.ones(n - 1)], [-1, 0, 1]) / 6
    dims = (6, 5, 4)
    hs = [1.0 / (n + 1) for n in dims]
    Ks, Ms = [k1(n, h) for n, h in zip(dims, hs)], [m1(n, h) for n, h in zip(dims, hs)]
    K = sp.kron(sp.kron(Ks[0], Ms[1]), Ms[2]) + sp.kron(sp.kron(Ms[0], Ks[1]), Ms[2]) + sp.kron(sp.kron(Ms[0], Ms[1]), Ks[2])
    Mass = sp.kron(sp.kron(Ms[0], Ms[1]), Ms[2])
    n = K.shape[0]
    D = sp.diags(np.exp(1j * np.random.default_rng(7).uniform(0, 2 * np.pi, n)))
    A = (D @ K @ D.conj()).tocsr()
    B = (D @ Mass @ D.conj()).tocsr()
    A, B = ((A + A.conj().T) * 0.5).tocsr(), ((B + B.conj().T) * 0.5).tocsr()
    w = np.sort(sla.eigh(A.toarray(), B.toarray(), eigvals_only=True))
    want, M0 = 5, 12
    assert w[want] - w[want - 1] > 1e-6 * w[want]
    Emin, Emax = 0.0, 0.5 * (w[want - 1] + w[want])
    Q0 = fo.seeded_subspace(n, M0)
    fpm = fo.feastinit()
    fpm[1] = 16
    r = fp.feast_hrr_mslanczos_gen_cheb(A, B, Emin, Emax, M0, fpm, Q0)
    ro = fo.feast_hcsrgv(A.tocsc(), B.tocsc(), Emin, Emax, M0, fo.feastinit(), Q0=Q0)
    assert r.info == ro.info == 0 and r.M == ro.M == want
    assert np.abs(np.sort(r.lambda_) - w[:want]).max() < 1e-10 * w[want] and r.res.max() < 1e-12
    assert fo.subspace_angle(np.asarray(r.q), np.asarray(ro.q, dtype=complex)) < 1e-8
    lo, hi = r.stats["cheb_interval"]
    ev = np.sort(sla.eigvalsh((sp.diags(1 / np.sqrt(B.diagonal().real)) @ B @ sp.diags(1 / np.sqrt(B.diagonal().real))).toarray()))
    assert lo <= ev[0] and hi >= ev[-1]          # the Chebyshev interval encloses the spectrum of D^-1 B


def test_two_sided_lanczos_port_on_the_reduced_config4_pencil():
    """The general filter AS THE ENGINE RUNS IT (two-sided multi-shift Lanczos, two passes, Jacobi sweeps for B): analytic eigenvalues
    of the Toeplitz-Kronecker pencil and the oracle's general solver."""
    dims = (8, 8, 6)
    coef = [(0.4 + 0.1j, 1.0 + 0.05j, 0.9 - 0.05j), (0.3 - 0.1j, 0.8 + 0.1j, 0.75 + 0.05j), (0.5 + 0.2j, 0.6 - 0.05j, 0.65 + 0.02j)]
    eps = 0.05
    toe = lambda n, a, b, c: sp.diags([b * np.ones(n - 1), a * np.ones(n), c * np.ones(n - 1)], [-1, 0, 1])
    I = [sp.identity(n) for n in dims]
    T = [toe(n, *abc) for n, abc in zip(dims, coef)]
    S = [toe(n, 0.0, abc[1], abc[2]) for n, abc in zip(dims, coef)]
    ksum = lambda X: sp.kron(sp.kron(X[0], I[1]), I[2]) + sp.kron(sp.kron(I[0], X[1]), I[2]) + sp.kron(sp.kron(I[0], I[1]), X[2])
    A = ksum(T).tocsr()
    B = (sp.identity(A.shape[0]) + eps * ksum(S)).tocsr()
    n = A.shape[0]
    la, ls = [], []
    for nn, (a, b, c) in zip(dims, coef):
        th = np.arange(1, nn + 1) * np.pi / (nn + 1)
        la.append(a + 2 * np.sqrt(b * c) * np.cos(th))
        ls.append(2 * np.sqrt(b * c) * np.cos(th))
    lamA = (la[0][:, None, None] + la[1][None, :, None] + la[2][None, None, :]).ravel()
    lamS = (ls[0][:, None, None] + ls[1][None, :, None] + ls[2][None, None, :]).ravel()
    lam = lamA / (1 + eps * lamS)
    order = np.argsort(lam.real)
    Emid = complex(lam[order[0]].real, lam[order[:17]].imag.mean())
    dist = np.abs(lam - Emid)
    rad = 0.5 * (np.sort(dist)[11] + np.sort(dist)[12])
    inside = lam[dist <= rad]
    assert len(inside) == 12
    M0 = 24
    Q0 = fo.seeded_subspace(n, M0)
    fpm = fo.feastinit()
    fpm[7], fpm[2] = 24, 10
    r = fp.feast_general_mstwosided(A, B, Emid, rad, M0, list(fpm), Q0)
    assert r.info == 0 and r.M == 12 and r.res.max() < 1e-10 and r.loop <= 3
    left = list(inside)
    for g in r.lambda_:
        j = int(np.argmin([abs(g - x) for x in left]))
        assert abs(g - left[j]) < 1e-9
        left.pop(j)
    ro = fo.feast_general(A.tocsc(), B.tocsc(), Emid, rad, M0, list(fpm), Q0=Q0, residual="true")
    assert ro.M == r.M
    ops = fp.jacobi_setup(A, B)
    assert 0 < ops[4] < 0.9           # Jacobi contraction bound of D^-1 B


def test_zolotarev_contour_tables():
    """fpm[16] = 2 (core/feast_tools.jl:44-210, 263-266): table nodes lie on the unit half circle, the rule is symmetric about the
    imaginary axis, and the rational filter it defines is the Zolotarev filter: ~1 inside the interval, decaying outside at the rate
    the reference quotes per table (n = 8: 1.12e-2 per FEAST loop)."""
    for ne in (1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 16, 20):
        fpm = fo.feastinit()
        fpm[1], fpm[15] = ne, 2
        Z, W = fo.feast_contour(-1.0, 1.0, fpm)
        assert len(Z) == ne and np.allclose(np.abs(Z), 1.0, atol=1e-2) and np.all(Z.imag > 0)    # (one node of FEAST's n = 20 table has |xe| = 0.9928)
        assert np.allclose(np.sort(Z.real), -np.sort(-Z.real)[::-1], atol=1e-15)
    fpm = fo.feastinit()
    fpm[1], fpm[15] = 8, 2
    Z, W = fo.feast_contour(2.0, 6.0, fpm)                       # mapped: Zne = xe * r + Emid, Wne = we * r
    xe, we = fo.zolotarev_point(8, 1)
    assert Z[0] == xe * 2.0 + 4.0 and W[0] == we * 2.0
    _, we0 = fo.zolotarev_point(8, 0)
    rho = lambda lam: (we0 + sum(2 * w / (z - lam) for z, w in zip(Z, W)).real) if False else sum(2 * (w / (z - lam)).real for z, w in zip(Z, W))
    inside = np.array([rho(x) for x in np.linspace(2.2, 5.8, 9)])
    outside = np.array([rho(x) for x in (0.0, 1.0, 7.0, 8.0, 40.0)])
    assert np.all(inside > 0.4) and np.abs(outside).max() < 0.1 * inside.min()
    # an ne without a table falls back like the reference does (core/feast_tools.jl:196-209)
    fpm = fo.feastinit()
    fpm[1], fpm[15] = 9, 2
    Z9, W9 = fo.feast_contour(-1.0, 1.0, fpm)
    assert np.allclose(W9, 1j * np.pi / 9)
