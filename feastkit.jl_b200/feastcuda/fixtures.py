"""Readers for the coordinate ("compact MatrixMarket-like") fixture files of the original FEAST example systems.

Mirror of the reader names in the reference's examples/feast/utils.jl:15-150 (``read_mm_dense_real`` … ``read_banded_complex``,
``read_polynomial_*``, ``to_complex_sparse``).  The format: a header line ``n  n  nnz`` followed by ``nnz`` lines
``i  j  value`` (real) or ``i  j  re  im`` (complex), 1-based indices, no MatrixMarket banner.  The ``system*.mtx`` files themselves are
not shipped with the reference, so every reader takes either a system NAME (looked up as ``<data_dir>/<name>.mtx``; ``data_dir`` defaults
to the ``FEASTCUDA_DATA_DIR`` environment variable) or a path / open text stream.  A standard MatrixMarket banner and ``%`` comment lines
in front of the header are skipped, so real ``.mtx`` files of the ``coordinate general`` kind load too.

The outputs are what the reference-named entry points of ``feastcuda.api`` take: dense column-major-agnostic ``numpy`` matrices,
``scipy.sparse.csc_matrix`` (the ``SparseMatrixCSC`` of the boundary, duplicates summed like Julia's ``sparse(row, col, val, n, n)``), and
LAPACK general-band arrays ``(kl + ku + 1) x n`` with the diagonal in row ``ku`` (0-based) together with ``(kl, ku)``.
Host-side only: nothing here touches the device.
"""
from __future__ import annotations

import io
import os

import numpy as np

__all__ = ["feast_data_path", "read_mm_dense_real", "read_mm_dense_complex", "read_mm_sparse_real", "read_mm_sparse_complex",
           "read_banded_real", "read_banded_complex", "read_polynomial_dense_real", "read_polynomial_sparse_real", "to_complex_sparse",
           "write_mm_coordinate"]


def feast_data_path(*parts, data_dir=None):
    """``feast_data_path(parts...)`` (utils.jl:13): join under the fixture directory (``data_dir`` or ``$FEASTCUDA_DATA_DIR``)."""
    base = data_dir if data_dir is not None else os.environ.get("FEASTCUDA_DATA_DIR")
    if base is None:
        raise FileNotFoundError("the FEAST example systems are not shipped: pass data_dir= or set FEASTCUDA_DATA_DIR")
    return os.path.join(base, *parts)


def _open(src, data_dir):
    if hasattr(src, "read"):
        return src, False
    path = os.fspath(src)
    if not os.path.exists(path):
        path = feast_data_path(path + ".mtx", data_dir=data_dir)
    return open(path, "r"), True


def _read_coordinates(src, complex_values, data_dir=None):
    """-> (n, rows0, cols0, values): one pass over the text, entries kept in file order (a later duplicate overwrites / adds downstream)."""
    fh, own = _open(src, data_dir)
    try:
        line = fh.readline()
        while line and (line.lstrip().startswith("%") or not line.strip()):
            line = fh.readline()
        head = line.split()
        if len(head) < 3:
            raise ValueError(f"fixture header must be 'n n nnz', got {line!r}")
        n, nnz = int(head[0]), int(head[2])
        if n <= 0 or nnz < 0:
            raise ValueError(f"fixture header with n = {n}, nnz = {nnz}")
        body = np.loadtxt(fh, dtype=np.float64, ndmin=2, max_rows=nnz) if nnz else np.zeros((0, 4 if complex_values else 3))
    finally:
        if own:
            fh.close()
    need = 4 if complex_values else 3
    if body.shape[0] != nnz or body.shape[1] < need:
        raise ValueError(f"fixture announces {nnz} entries of {need} fields, found an array of shape {body.shape}")
    rows = body[:, 0].astype(np.int64) - 1
    cols = body[:, 1].astype(np.int64) - 1
    if nnz and (rows.min() < 0 or cols.min() < 0 or rows.max() >= n or cols.max() >= n):
        raise ValueError("fixture entry outside the n x n matrix (indices are 1-based)")
    vals = body[:, 2] + 1j * body[:, 3] if complex_values else body[:, 2].copy()
    return n, rows, cols, vals


def _dense(src, cplx, data_dir):
    n, r, c, v = _read_coordinates(src, cplx, data_dir)
    M = np.zeros((n, n), dtype=np.complex128 if cplx else np.float64, order="F")
    M[r, c] = v                     # assignment like the reference: the last duplicate wins
    return M


def _sparse(src, cplx, data_dir):
    import scipy.sparse as sp
    n, r, c, v = _read_coordinates(src, cplx, data_dir)
    M = sp.coo_matrix((v, (r, c)), shape=(n, n)).tocsc()      # duplicates are summed, like sparse(row, col, val, n, n)
    M.sort_indices()
    return M


def _banded(src, cplx, data_dir):
    n, r, c, v = _read_coordinates(src, cplx, data_dir)
    kl = int(max(0, (r - c).max())) if r.size else 0
    ku = int(max(0, (c - r).max())) if r.size else 0
    band = np.zeros((kl + ku + 1, n), dtype=np.complex128 if cplx else np.float64, order="F")
    band[ku + r - c, c] = v
    return band, kl, ku


def read_mm_dense_real(name, data_dir=None):
    """utils.jl:15-31 -> ``Matrix{Float64}``."""
    return _dense(name, False, data_dir)


def read_mm_dense_complex(name, data_dir=None):
    """utils.jl:33-50 -> ``Matrix{ComplexF64}``."""
    return _dense(name, True, data_dir)


def read_mm_sparse_real(name, data_dir=None):
    """utils.jl:52-69 -> ``SparseMatrixCSC{Float64,Int}`` (here ``scipy.sparse.csc_matrix``)."""
    return _sparse(name, False, data_dir)


def read_mm_sparse_complex(name, data_dir=None):
    """utils.jl:71-90 -> ``SparseMatrixCSC{ComplexF64,Int}``."""
    return _sparse(name, True, data_dir)


def read_banded_real(name, data_dir=None):
    """utils.jl:92-121 -> ``(band, k_lower, k_upper)``: LAPACK general band storage, diagonal in row ``k_upper`` (0-based)."""
    return _banded(name, False, data_dir)


def read_banded_complex(name, data_dir=None):
    """utils.jl:123-154 -> ``(band, k_lower, k_upper)`` with complex entries."""
    return _banded(name, True, data_dir)


def read_polynomial_dense_real(prefix, data_dir=None):
    """utils.jl:156-162: the three coefficient matrices ``<prefix>A0, A1, A2`` of a quadratic eigenproblem, dense."""
    return [read_mm_dense_real(f"{prefix}A{k}", data_dir) for k in range(3)]


def read_polynomial_sparse_real(prefix, data_dir=None):
    """utils.jl:164-170: the same, sparse."""
    return [read_mm_sparse_real(f"{prefix}A{k}", data_dir) for k in range(3)]


def to_complex_sparse(A):
    """utils.jl:172-174: same pattern, ``ComplexF64`` values."""
    return A.astype(np.complex128)


def write_mm_coordinate(dst, M, complex_values=None):
    """Write a dense or sparse matrix in the fixture format (header ``n n nnz``; used by the tests to build fixtures, and the way to
    hand a matrix to the original FEAST example drivers).  Entries are emitted column by column."""
    import scipy.sparse as sp
    C = sp.coo_matrix(M.tocsc() if sp.issparse(M) else sp.csc_matrix(np.asarray(M)))
    if complex_values is None:
        complex_values = np.iscomplexobj(C.data)
    out = io.StringIO()
    out.write(f"{C.shape[0]} {C.shape[1]} {C.nnz}\n")
    for i, j, v in zip(C.row, C.col, C.data):
        out.write(f"{i + 1} {j + 1} {float(np.real(v))!r} {float(np.imag(v))!r}\n" if complex_values else f"{i + 1} {j + 1} {float(np.real(v))!r}\n")
    text = out.getvalue()
    if hasattr(dst, "write"):
        dst.write(text)
    else:
        with open(os.fspath(dst), "w") as fh:
            fh.write(text)
