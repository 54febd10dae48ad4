"""GPU parity tests of the matrix-free interface: feast_matvec / feast_sparse_matvec! (interfaces/feast_interfaces.jl:465-481,
sparse/feast_sparse.jl:1284-1471) with DEVICE operators -- a compiled CUDA callback (examples/matfree_laplacian.cu) and a Python
callable working on torch views of the library's blocks.  Eigenpairs against the oracle run on the assembled matrix."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

import feast_oracle as fo

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


class LaplacianGrid(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int)]


def _check(r, ro, A):
    assert r.info == ro.info == 0 and r.M == ro.M
    assert np.abs(np.sort(r.lambda_) - np.sort(ro.lambda_)).max() <= 1e-10 * max(1.0, np.abs(ro.lambda_).max())
    assert r.res.max() < 1e-12
    assert fo.subspace_angle(np.asarray(r.q, dtype=complex), np.asarray(ro.q, dtype=complex)) < 1e-8
    for j in range(r.M):
        assert np.linalg.norm(A @ r.q[:, j] - r.lambda_[j] * r.q[:, j]) < 1e-11 * max(1.0, abs(r.lambda_[j]))


@pytest.mark.parametrize("N,M0,k_in", [(16, 24, 9), (14, 7, 3)])
def test_feast_matvec_with_a_compiled_cuda_operator(N, M0, k_in):
    """The user's operator is a CUDA kernel launcher (7-point stencil, nothing assembled); the solve must find the pairs of the
    assembled Laplacian, with the same sweeps and nearly the same Lanczos steps as the CSR path (same recurrence, the mat-vec's
    summation order differs)."""
    import feastcuda as fc
    ex = C.CDLL(str(ROOT / "feastkit.jl_b200" / "lib" / "libfeastcuda_examples.so"))
    grid = LaplacianGrid(N, N, N)
    A = fo.laplacian_3d(N).astype(float).tocsc()
    ev = fo.laplacian_3d_eigs(N)
    gaps = [i for i in range(k_in, k_in + 30) if ev[i + 1] - ev[i] > 1e-3]
    Emin, Emax = 0.0, 0.5 * (ev[gaps[0]] + ev[gaps[0] + 1])
    Q0 = fo.seeded_subspace(N ** 3, M0, complex_storage=False)
    ro = fo.feast_scsrev(A, Emin, Emax, M0, fo.feastinit(), Q0=Q0.astype(complex), filter="true")
    r = fc.feast_matvec((ex.feastcuda_example_laplacian3d, grid), None, N ** 3, (Emin, Emax), M0=M0, fpm=fc.feastinit(), Q0=Q0,
                        solver_maxiter=2000)
    _check(r, ro, A)
    rs = fc.feast_scsrev(A, Emin, Emax, M0, fc.feastinit(), Q0=Q0, solver_maxiter=2000, mixed=False)    # matrix-free runs FP64 vectors
    assert r.loop == rs.loop and abs(r.stats["lz_steps_p1"] - rs.stats["lz_steps_p1"]) <= 16
    assert r.stats["lz_steps_p2"] == r.stats["lz_steps_p1"] > 0


def test_feast_matvec_with_python_device_callables():
    """A_mul!(y, x) / B_mul!(y, x) as Python callables on torch CUDA views (1-D Laplacian by shifted slices, B = identity
    copy); a non-identity B is refused, an exception inside the callback surfaces as that exception."""
    import torch
    import feastcuda as fc
    n, M0 = 400, 12
    A = fo.laplacian_1d(n).tocsc()
    w = np.linalg.eigvalsh(A.toarray())
    Emin, Emax = 0.0, 0.5 * (w[5] + w[6])
    calls = {"n": 0}

    def A_mul(Y, X):
        calls["n"] += 1
        Y.copy_(2.0 * X)
        Y[1:] -= X[:-1]
        Y[:-1] -= X[1:]

    def B_mul(Y, X):
        Y.copy_(X)

    Q0 = fo.seeded_subspace(n, M0, complex_storage=False)
    ro = fo.feast_scsrev(A, Emin, Emax, M0, fo.feastinit(), Q0=Q0.astype(complex), filter="true")
    r = fc.feast_matvec(A_mul, B_mul, n, (Emin, Emax), M0=M0, fpm=fc.feastinit(), Q0=Q0, solver_maxiter=4000)
    _check(r, ro, A)
    assert r.M == 6 and calls["n"] > r.stats["lz_steps_p1"]
    assert r.q.dtype == np.float64

    with pytest.raises(NotImplementedError):
        fc.feast_matvec(A_mul, lambda Y, X: Y.copy_(2.0 * X), n, (Emin, Emax), M0=M0, fpm=fc.feastinit(), Q0=Q0)

    def broken(Y, X):
        raise ZeroDivisionError("operator failed")

    with pytest.raises(ZeroDivisionError):
        fc.feast_matvec(broken, None, n, (Emin, Emax), M0=M0, fpm=fc.feastinit(), Q0=Q0)
    # the engine is usable again afterwards
    r2 = fc.feast_matvec(A_mul, None, n, (Emin, Emax), M0=M0, fpm=fc.feastinit(), Q0=Q0, solver_maxiter=4000)
    assert r2.info == 0 and r2.M == 6
    assert torch.cuda.is_available()


def test_matrix_free_operators_refuse_what_they_cannot_serve(engine):
    """B != I, complex subspaces and the general contour need solvers the matrix-free path does not have: UNSUPPORTED, not a
    fallback."""
    import feastcuda as fc
    n = 64

    def A_mul(Y, X):
        Y.copy_(3.0 * X)

    engine.set_matfree(n, A_mul)
    engine.clear_b()
    fpm = fc.feastinit()
    fc.feastdefault_(fpm)
    Z, W = fc.feast_gcontour(3.0 + 0j, 1.0, fpm)
    with pytest.raises(fc.FeastCudaError) as ei:
        engine.solve_contour(3.0 + 0j, 1.0, 4, fpm, Z, W, Q0=fo.seeded_subspace(n, 4))
    assert ei.value.code == fc._lib.ERR_UNSUPPORTED
