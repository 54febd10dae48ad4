"""CPU tests of the fixture readers (feastcuda/fixtures.py): the coordinate format of the original FEAST example systems as the reference's
examples/feast/utils.jl:15-150 reads it (header `n n nnz`, 1-based `i j value` / `i j re im` lines)."""
import io
import textwrap

import numpy as np
import pytest
import scipy.sparse as sp

import feast_oracle as fo

REAL = textwrap.dedent("""\
    4 4 7
    1 1 2.0
    2 2 3.0
    3 3 4.0
    4 4 5.0
    1 2 -1.0
    2 1 -1.0
    4 2 0.5
    """)
CPLX = textwrap.dedent("""\
    3 3 5
    1 1 2.0 0.0
    2 2 3.0 0.0
    3 3 4.0 0.0
    1 2 0.1 0.2
    2 1 0.1 -0.2
    """)


def test_dense_sparse_and_banded_readers_agree_on_hand_written_fixtures():
    import feastcuda as fc
    want = np.array([[2, -1, 0, 0], [-1, 3, 0, 0], [0, 0, 4, 0], [0, 0.5, 0, 5.0]])
    D = fc.read_mm_dense_real(io.StringIO(REAL))
    S = fc.read_mm_sparse_real(io.StringIO(REAL))
    assert D.dtype == np.float64 and np.array_equal(D, want)
    assert sp.isspmatrix_csc(S) and S.nnz == 7 and np.array_equal(S.toarray(), want) and S.has_sorted_indices
    band, kl, ku = fc.read_banded_real(io.StringIO(REAL))
    assert (kl, ku) == (2, 1) and band.shape == (4, 4)
    for i in range(4):                      # LAPACK general band: entry (i, j) sits in row ku + i - j of column j
        for j in range(4):
            if -ku <= i - j <= kl:
                assert band[ku + i - j, j] == want[i, j]
    Z = fc.read_mm_dense_complex(io.StringIO(CPLX))
    assert Z.dtype == np.complex128 and Z[0, 1] == 0.1 + 0.2j and Z[1, 0] == 0.1 - 0.2j and np.allclose(Z, Z.conj().T)
    Zs = fc.read_mm_sparse_complex(io.StringIO(CPLX))
    assert np.array_equal(Zs.toarray(), Z)
    zb, kl, ku = fc.read_banded_complex(io.StringIO(CPLX))
    assert (kl, ku) == (1, 1) and zb[ku + 1 - 0, 0] == 0.1 - 0.2j and zb[ku + 0 - 1, 1] == 0.1 + 0.2j
    assert fc.to_complex_sparse(S).dtype == np.complex128 and np.array_equal(fc.to_complex_sparse(S).toarray().real, want)


def test_named_systems_round_trip_through_files(tmp_path, monkeypatch):
    import feastcuda as fc
    rng = np.random.default_rng(3)
    A = fo.laplacian_1d(12).tocsc()
    G = sp.random(9, 9, density=0.4, random_state=5).tocsc() + 1j * sp.random(9, 9, density=0.4, random_state=6).tocsc()
    fc.write_mm_coordinate(tmp_path / "system1.mtx", A)
    fc.write_mm_coordinate(tmp_path / "system4.mtx", G)
    for k in range(3):
        fc.write_mm_coordinate(tmp_path / f"system5A{k}.mtx", np.diag(rng.standard_normal(5)) + (k == 1) * np.eye(5, k=1))
    # by name under data_dir=, by name under $FEASTCUDA_DATA_DIR, by path
    assert np.array_equal(fc.read_mm_sparse_real("system1", data_dir=tmp_path).toarray(), A.toarray())
    monkeypatch.setenv("FEASTCUDA_DATA_DIR", str(tmp_path))
    assert fc.feast_data_path("system1.mtx") == str(tmp_path / "system1.mtx")
    assert np.array_equal(fc.read_mm_dense_real("system1"), A.toarray())
    assert np.array_equal(fc.read_mm_sparse_complex(tmp_path / "system4.mtx").toarray(), G.toarray())
    band, kl, ku = fc.read_banded_real("system1")
    assert (kl, ku) == (1, 1) and np.array_equal(band[1], A.diagonal()) and np.array_equal(band[0, 1:], A.diagonal(1))
    P = fc.read_polynomial_dense_real("system5")
    Ps = fc.read_polynomial_sparse_real("system5")
    assert len(P) == len(Ps) == 3 and all(np.array_equal(a, b.toarray()) for a, b in zip(P, Ps)) and P[1][0, 1] == 1.0
    monkeypatch.delenv("FEASTCUDA_DATA_DIR")
    with pytest.raises(FileNotFoundError):
        fc.read_mm_dense_real("system1")


def test_banner_duplicates_and_malformed_files():
    import feastcuda as fc
    mm = "%%MatrixMarket matrix coordinate real general\n% a comment\n" + REAL
    assert np.array_equal(fc.read_mm_dense_real(io.StringIO(mm)), fc.read_mm_dense_real(io.StringIO(REAL)))
    dup = "2 2 3\n1 1 1.0\n1 1 2.5\n2 2 1.0\n"
    assert fc.read_mm_dense_real(io.StringIO(dup))[0, 0] == 2.5             # A[i, j] = val: the last one wins (utils.jl:26)
    assert fc.read_mm_sparse_real(io.StringIO(dup))[0, 0] == 3.5            # sparse(row, col, val): duplicates add (utils.jl:68)
    empty = fc.read_mm_sparse_real(io.StringIO("3 3 0\n"))
    assert empty.shape == (3, 3) and empty.nnz == 0
    for bad in ("2 2\n", "2 2 2\n1 1 1.0\n", "2 2 1\n3 1 1.0\n", "2 2 1\n0 1 1.0\n"):
        with pytest.raises(ValueError):
            fc.read_mm_dense_real(io.StringIO(bad))
    with pytest.raises(ValueError):
        fc.read_mm_dense_complex(io.StringIO(REAL))                         # three fields where four are needed
