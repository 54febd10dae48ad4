"""CPU test of the band-kernel restatement (oracle/feast_port.py band_shift / band_lu / band_solve_window = csrc/kernels_band.cuh
k_band_shift / k_band_lu_warp / k_band_solve_win) against LAPACK's zgbtrf / zgbtrs (banded/feast_banded.jl:108,141,678,683)."""
import numpy as np
import pytest
import scipy.linalg as sla

import feast_port as fp


def _general_band(rng, n, k):
    AB = np.zeros((2 * k + 1, n), dtype=complex)
    Af = np.zeros((n, n), dtype=complex)
    for j in range(n):
        for i in range(max(0, j - k), min(n - 1, j + k) + 1):
            AB[k + i - j, j] = Af[i, j] = rng.standard_normal() + 1j * rng.standard_normal()
    return AB, Af


@pytest.mark.parametrize("n,ka,kb,K", [(1, 0, 0, 2), (2, 1, 0, 2), (5, 1, 1, 2), (9, 3, 2, 4), (40, 7, 2, 8), (33, 7, 7, 16), (20, 2, 5, 8),
                                       (12, 4, 0, 4), (70, 9, 3, 16)])
def test_band_factor_layout_pivots_and_window_solve_match_lapack(n, ka, kb, K):
    rng = np.random.default_rng(100 * n + ka)
    AB, Af = _general_band(rng, n, ka)
    BB, Bf = _general_band(rng, n, kb)
    z = 0.3 + 0.2j
    for with_b in (True, False):
        F, k = fp.band_shift(AB, BB if with_b else None, z, ka, kb)
        S = z * (Bf if with_b else np.eye(n)) - Af
        assert k == (max(ka, kb) if with_b else ka) and F.shape == (3 * k + 1, n)
        assert not F[:k].any()                                               # fill-in rows start at zero
        for j in range(n):
            for i in range(max(0, j - k), min(n - 1, j + k) + 1):
                assert abs(F[2 * k + i - j, j] - S[i, j]) < 1e-14
        lu, piv, info_l = sla.lapack.zgbtrf(F.copy(), k, k)
        ipiv, info = fp.band_lu(F, k)
        assert info == info_l == 0
        assert np.array_equal(ipiv, piv) or np.array_equal(ipiv + 1, piv)    # same pivot rows as LAPACK (0- or 1-based wrapper)
        assert np.abs(lu - F).max() < 1e-12 * max(1.0, np.abs(lu).max())
        for Kw in sorted({max(K, k), max(2 * K, k)}):                        # a window wider than the band changes nothing
            b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
            x = fp.band_solve_window(F, ipiv, k, Kw, b)
            want = np.linalg.solve(S, b)
            assert np.abs(x - want).max() < 1e-10 * np.abs(want).max()


def test_band_lu_reports_the_first_zero_pivot_like_zgbtf2():
    AB = np.zeros((3, 4), dtype=complex)
    AB[1] = [1.0, 0.0, 2.0, 3.0]          # diagonal with a zero, no off-diagonals: z = 0 -> -A singular in column 2
    F, k = fp.band_shift(AB, None, 0.0, 1, 0)
    lu, piv, info_l = sla.lapack.zgbtrf(F.copy(), 1, 1)
    _, info = fp.band_lu(F, k)
    assert info == info_l == 2


def test_oracle_banded_route_equals_its_sparse_route():
    """feast_oracle.feast_hrr(band_k=...) -- per-node LAPACK band LUs, the reference's banded route (banded/feast_banded.jl:561-823) --
    returns the eigenpairs of the SuperLU route on a banded pencil (KA8-like operator at n = 60, k = 3, generalized)."""
    import scipy.sparse as sp
    import feast_oracle as fo
    rng = np.random.default_rng(5)
    n, k = 60, 3
    A = sp.diags([rng.standard_normal(n - abs(d)) * (2.0 if d == 0 else 0.4) for d in range(k + 1)], list(range(k + 1)))
    A = (A + sp.triu(A, 1).T).tocsc()
    B = sp.diags([rng.uniform(1.0, 2.0, n), rng.uniform(0.05, 0.1, n - 1)], [0, 1])
    B = (B + sp.triu(B, 1).T).tocsc()
    w = sla.eigh(A.toarray(), B.toarray(), eigvals_only=True)
    Emin, Emax = 0.5 * (w[19] + w[20]), 0.5 * (w[27] + w[28])
    Q0 = fo.seeded_subspace(n, 16, complex_storage=False).astype(complex)
    r1 = fo.feast_hrr(A, B, Emin, Emax, 16, fo.feastinit(), Q0=Q0, filter="true")
    r2 = fo.feast_hrr(A, B, Emin, Emax, 16, fo.feastinit(), Q0=Q0, filter="true", band_k=k)
    assert r1.info == r2.info == 0 and r1.M == r2.M == 8 and r1.loop == r2.loop
    assert np.abs(np.sort(r1.lambda_) - w[20:28]).max() < 1e-12 and np.abs(np.sort(r2.lambda_) - w[20:28]).max() < 1e-12
    assert fo.subspace_angle(r1.q, r2.q) < 1e-9


def test_oracle_node_pool_equals_the_sequential_node_loop():
    """feast_oracle.feast_hrr(node_pool=...) -- one worker per quadrature node with its cached factorisation, the shape of the reference's
    :threads backend (parallel/feast_parallel.jl:586) -- returns bit-identical pairs to the sequential node loop (:serial)."""
    from concurrent.futures import ThreadPoolExecutor
    import feast_oracle as fo
    A = fo.laplacian_3d(7).tocsc().astype(complex)
    ev = fo.laplacian_3d_eigs(7)
    Emin, Emax = 0.0, 0.5 * (ev[9] + ev[10])
    Q0 = fo.seeded_subspace(343, 20, complex_storage=False).astype(complex)
    fpm = fo.feastinit()
    fo.feastdefault(fpm)
    Zne, Wne = fo.feast_contour(Emin, Emax, fpm)
    facs = [None] * len(Zne)

    def one(e, rhs):
        if facs[e] is None:
            facs[e] = fo._factor(A, None, Zne[e])
        return 2 * Wne[e] * fo._solve_factor(facs[e], rhs)

    def pool(rhs):
        with ThreadPoolExecutor(len(Zne)) as ex:
            return sum(ex.map(lambda e: one(e, rhs), range(len(Zne))))
    r1 = fo.feast_hrr(A, None, Emin, Emax, 20, fo.feastinit(), Q0=Q0, filter="true")
    r2 = fo.feast_hrr(A, None, Emin, Emax, 20, fo.feastinit(), Q0=Q0, filter="true", node_pool=pool)
    assert r1.info == r2.info == 0 and r1.M == r2.M == 10 and r1.loop == r2.loop
    assert np.abs(np.sort(r1.lambda_) - np.sort(r2.lambda_)).max() < 1e-13 and np.abs(np.sort(r2.lambda_) - ev[:10]).max() < 1e-12
