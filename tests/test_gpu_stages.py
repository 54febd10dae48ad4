"""GPU parity tests of the stage-level C-ABI entry points against the CPU oracle (same seeded inputs).

Tolerances: everything here is Float64/ComplexF64 arithmetic; reductions are re-associated on the GPU, so
stage outputs are compared at 1e-12 relative (1e-10 for solves, BASELINE.json's eigenvalue tolerance).
"""
import numpy as np
import pytest
import scipy.linalg as sla
import scipy.sparse as sp

import feast_oracle as fo
import feast_port as fp

pytestmark = pytest.mark.gpu


def _rand_block(rng, n, m):
    return rng.standard_normal((n, m)) + 1j * rng.standard_normal((n, m))


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.fixture()
def lap(engine):
    import feastcuda as fc
    A = fo.laplacian_3d(9, 8, 7).astype(np.float64)
    engine.set_sparse(fc.A, A.tocsc(), fc.SYM)
    engine.clear_b()
    return A


@pytest.mark.parametrize("m", [1, 3, 4, 8, 13, 16, 32, 33, 64, 96, 100, 128])
def test_spmm_shifted_real_identity_b(engine, lap, m):
    rng = np.random.default_rng(m)
    X = _rand_block(rng, lap.shape[0], m)
    z = 0.3 + 0.05j
    Y = engine.spmm_shifted(z, X)
    ref = fp.shifted_apply(lap, None, z, X)
    assert _rel(Y, ref) < 1e-14


def test_spmm_ragged_rows_and_empty_rows(engine):
    import feastcuda as fc
    rng = np.random.default_rng(5)
    n = 301
    M = sp.random(n, n, density=0.05, random_state=3, format="csr")
    M = (M + M.T).tolil()
    M[17, :] = 0
    M[:, 17] = 0  # an empty row/column
    M[5, :] = rng.standard_normal(n)  # a dense row (> 32 entries: several shuffle rounds)
    M[:, 5] = M[5, :].T
    M = M.tocsr()
    engine.set_sparse(fc.A, M.tocsc(), fc.SYM)
    engine.clear_b()
    for m in (2, 8, 40):
        X = _rand_block(rng, n, m)
        z = -0.7 + 0.4j
        assert _rel(engine.spmm_shifted(z, X), fp.shifted_apply(M, None, z, X)) < 1e-14
        assert _rel(engine.apply(fc.A, X), M @ X) < 1e-14


def test_spmm_complex_hermitian_with_b_csc_conjugation(engine):
    """SparseMatrixCSC of a Hermitian matrix read as CSR is conj(A): the library conjugates on load."""
    import feastcuda as fc
    rng = np.random.default_rng(11)
    n = 257
    R = sp.random(n, n, density=0.03, random_state=1, format="csr")
    I = sp.random(n, n, density=0.03, random_state=2, format="csr")
    Ah = (R + R.T) + 1j * (I - I.T)
    Bh = sp.identity(n) * 2.0 + 0.1 * ((R + R.T) + 1j * (I - I.T))
    engine.set_sparse(fc.A, sp.csc_matrix(Ah), fc.HERM)
    engine.set_sparse(fc.B, sp.csc_matrix(Bh), fc.HERM)
    X = _rand_block(rng, n, 24)
    z = 1.5 - 0.2j
    assert _rel(engine.spmm_shifted(z, X), fp.shifted_apply(Ah.tocsr(), Bh.tocsr(), z, X)) < 1e-14
    assert _rel(engine.apply(fc.B, X), Bh @ X) < 1e-14
    engine.clear_b()


def test_spmm_general_csc_is_transposed_once(engine):
    import feastcuda as fc
    rng = np.random.default_rng(12)
    n = 120
    G = sp.random(n, n, density=0.08, random_state=7, format="csc") + 1j * sp.random(n, n, density=0.08, random_state=8, format="csc")
    engine.set_sparse(fc.A, sp.csc_matrix(G), fc.GEN)
    engine.clear_b()
    X = _rand_block(rng, n, 7)
    assert _rel(engine.apply(fc.A, X), G @ X) < 1e-14


@pytest.mark.parametrize("n,m", [(50, 1), (1000, 7), (4097, 64), (300, 128), (70000, 33)])
def test_gram(engine, n, m):
    rng = np.random.default_rng(n + m)
    X = _rand_block(rng, n, m)
    Y = _rand_block(rng, n, m)
    G = engine.gram(X, Y)
    assert _rel(G, X.conj().T @ Y) < 1e-13


def test_accumulate(engine, lap):
    rng = np.random.default_rng(3)
    n = lap.shape[0]
    Y = _rand_block(rng, n, 20)
    Q = _rand_block(rng, n, 20)
    w = 0.3 - 1.1j
    assert _rel(engine.accumulate(w, Y, Q), Q + w * Y) < 1e-15


def test_orthonormalize_reference_rank2_case(engine):
    """test/test_allocation_helpers.jl:274-292: rank-2 4x4 block, Q'Q = I, span preserved."""
    src = np.array([[1, 2, 0, 1e-15], [1j, 2j, 1, 1e-15j], [0, 0, 1j, 0], [0, 0, 0, 0]], dtype=complex)
    Q, rank = engine.orthonormalize(src)
    Qo, ranko = fo.qr_compress(src, 4)
    assert rank == ranko == 2
    assert np.allclose(Q.conj().T @ Q, np.eye(rank), atol=1e-12)
    assert np.linalg.norm(src - Q @ (Q.conj().T @ src)) <= 1e-12


@pytest.mark.parametrize("n,m,true_rank,decay", [(500, 32, 32, 1.0), (2000, 64, 40, 1.0), (3000, 64, 64, 1e-7),
                                                  (800, 16, 5, 1.0), (5000, 128, 100, 1e-5)])
def test_orthonormalize_rank_and_span_match_pivoted_qr(engine, n, m, true_rank, decay):
    rng = np.random.default_rng(n)
    U, _ = np.linalg.qr(_rand_block(rng, n, true_rank))
    s = np.logspace(0, np.log10(decay), true_rank)
    W = (U * s) @ _rand_block(rng, true_rank, m)
    Q, rank = engine.orthonormalize(W)
    Qo, ranko = fo.qr_compress(W, m)
    assert rank == ranko
    assert np.abs(Q.conj().T @ Q - np.eye(rank)).max() < 1e-13
    # same subspace as the pivoted-QR basis
    assert np.linalg.norm(Qo - Q @ (Q.conj().T @ Qo), 2) < 1e-7 * max(1.0, 1.0 / (decay * 1e6)) + 1e-9


@pytest.mark.parametrize("r", [1, 2, 5, 16, 33, 64, 97, 128])
def test_reduced_eig_standard(engine, r):
    rng = np.random.default_rng(r)
    S = _rand_block(rng, r, r)
    S = (S + S.conj().T) / 2
    lam, V, sweeps = engine.reduced_eig(S)
    w = np.linalg.eigvalsh(S)
    assert np.abs(lam - w).max() < 1e-12 * max(1.0, np.abs(w).max())
    assert np.abs(V.conj().T @ V - np.eye(r)).max() < 1e-12
    assert np.abs(S @ V - V * lam).max() < 1e-11 * max(1.0, np.abs(w).max())


@pytest.mark.parametrize("r", [1, 3, 20, 64, 128])
def test_reduced_eig_generalized_and_real_input(engine, r):
    rng = np.random.default_rng(100 + r)
    S = _rand_block(rng, r, r)
    S = (S + S.conj().T) / 2
    Bm = _rand_block(rng, r, r)
    Bm = Bm @ Bm.conj().T + r * np.eye(r)
    lam, V, _ = engine.reduced_eig(S, Bm)
    w = sla.eigh(S, Bm, eigvals_only=True)
    assert np.abs(lam - w).max() < 1e-11
    assert np.abs(V.conj().T @ Bm @ V - np.eye(r)).max() < 1e-11
    assert np.abs(S @ V - (Bm @ V) * lam).max() < 1e-10 * np.abs(Bm).max()
    # real symmetric input in complex storage gives real vectors (needed by the real-valued filter mode)
    Sr = S.real
    lam2, V2, _ = engine.reduced_eig(Sr)
    assert np.abs(V2.imag).max() == 0.0
    assert np.abs(lam2 - np.linalg.eigvalsh(Sr)).max() < 1e-12 * max(1.0, np.abs(lam2).max())


def test_residuals_match_reference_helper(engine):
    """test/test_allocation_helpers.jl:183-209 with the 3x3 pencil of that test."""
    import feastcuda as fc
    A = np.array([[4.0, 0.2, 0.0], [0.2, 5.0, 0.3], [0.0, 0.3, 6.0]])
    Bd = np.diag([1.0, 1.2, 1.5])
    q = np.array([[0.8, 0.1], [0.3, 0.7], [0.5, 0.6]])
    lam = np.array([4.2, 5.8])
    engine.set_sparse(fc.A, sp.csc_matrix(A), fc.SYM)
    engine.set_sparse(fc.B, sp.csc_matrix(Bd), fc.SYM)
    res = engine.residuals(q, lam)
    assert np.allclose(res, fo.feast_residual(A, Bd, lam, q, 2), rtol=1e-13)
    engine.clear_b()


@pytest.mark.parametrize("m", [1, 5, 32, 64])
def test_block_solve_matches_direct_solution(engine, m):
    import feastcuda as fc
    A = fo.laplacian_3d(8, 7, 6).astype(np.float64)
    n = A.shape[0]
    engine.set_sparse(fc.A, A.tocsc(), fc.SYM)
    engine.clear_b()
    rng = np.random.default_rng(m)
    RHS = _rand_block(rng, n, m)
    z = 0.35 + 0.08j
    X, iters, resid = engine.block_solve(z, RHS, solver_tol=1e-13, solver_maxiter=1500)
    Xd = np.linalg.solve(z * np.eye(n) - A.toarray(), RHS)
    assert _rel(X, Xd) < 1e-10
    assert resid.max() <= 10 * 1e-13 * (1 + np.linalg.norm(RHS, axis=0).max())
    # same lock-step recurrence as the CPU port: away from the round-off floor the iteration counts agree
    _, it10, _ = engine.block_solve(z, RHS, solver_tol=1e-9, solver_maxiter=1500)
    _, its_p, _, _ = fp.block_bicgstab(A, None, z, RHS, rtol=1e-9, maxiter=1500)
    assert np.abs(it10 - its_p).max() <= max(3, 0.15 * its_p.max())


def test_block_solve_initial_guess_and_inexact_stop(engine):
    import feastcuda as fc
    A = fo.laplacian_3d(8).astype(np.float64)
    n = A.shape[0]
    engine.set_sparse(fc.A, A.tocsc(), fc.SYM)
    rng = np.random.default_rng(9)
    RHS = _rand_block(rng, n, 16)
    z = 0.5 + 0.1j
    Xd = np.linalg.solve(z * np.eye(n) - A.toarray(), RHS)
    X0 = Xd + 1e-6 * _rand_block(rng, n, 16)
    X, iters, resid = engine.block_solve(z, RHS, X0=X0, solver_tol=1e-13, solver_maxiter=1000, inner_rel=1e-2)
    r0 = np.linalg.norm(RHS - fp.shifted_apply(A, None, z, X0), axis=0)
    assert (resid <= 10 * np.maximum(1e-2 * r0, 1e-13 * (1 + np.linalg.norm(RHS, axis=0)))).all()
    assert iters.max() < 60  # two digits only
    Xg, it_g, _ = engine.block_solve(z, RHS, X0=Xd, solver_tol=1e-10, solver_maxiter=1000)
    assert it_g.max() == 0 and _rel(Xg, Xd) < 1e-13  # exact guess: zero iterations, untouched


def test_block_solve_generalized_hermitian(engine):
    import feastcuda as fc
    n = 200
    rng = np.random.default_rng(21)
    T = sp.diags([-np.ones(n - 1), 2 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1])
    D = sp.diags(np.exp(1j * rng.uniform(0, 2 * np.pi, n)))
    Ah = (D @ T @ D.conj().T).tocsc()
    Bh = (D @ sp.diags([np.ones(n - 1) / 6, 4 * np.ones(n) / 6, np.ones(n - 1) / 6], [-1, 0, 1]) @ D.conj().T).tocsc()
    engine.set_sparse(fc.A, Ah, fc.HERM)
    engine.set_sparse(fc.B, Bh, fc.HERM)
    RHS = _rand_block(rng, n, 12)
    z = 0.8 + 0.3j
    X, iters, resid = engine.block_solve(z, RHS, solver_tol=1e-12, solver_maxiter=3000)
    Xd = np.linalg.solve(z * Bh.toarray() - Ah.toarray(), RHS)
    assert _rel(X, Xd) < 1e-9
    engine.clear_b()
