"""Host-side passes over dense operators on their way to the C ABI (symmetry checks, layout copies), tiled and spread over the host
cores: at configs[1] (n = 8192) the plain NumPy forms of these two passes cost 3.8 s of a 4.8 s end-to-end solve."""
from __future__ import annotations

import os

import numpy as np

_TILE = 512          # tile edge of the blocked host passes over dense operators (a tile pair stays in L2 of the host cores)


def _tile_pool():
    from concurrent.futures import ThreadPoolExecutor
    return ThreadPoolExecutor(max(1, min(16, os.cpu_count() or 1)))


def dense_equals_own_adjoint(M, conj):
    """M == M^T (or M^H) for a square dense matrix.  Tile pairs (i, j) / (j, i) are compared by a thread pool (NumPy releases the GIL in
    the comparison): a strided full-matrix `M == M.T` walks one operand against the cache lines, 1.9 s at n = 8192 against 0.1 s here --
    that check was 40 % of the end-to-end time of configs[1]."""
    n = M.shape[0]
    conj = conj and np.iscomplexobj(M)
    if n <= 2 * _TILE:
        return bool(np.array_equal(M, M.conj().T if conj else M.T))

    def same(t):
        i, j = t
        other = M[j:j + _TILE, i:i + _TILE].T
        return bool(np.array_equal(M[i:i + _TILE, j:j + _TILE], other.conj() if conj else other))
    tiles = [(i, j) for i in range(0, n, _TILE) for j in range(i, n, _TILE)]
    with _tile_pool() as ex:
        return all(ex.map(same, tiles))


def column_major(M, dtype):
    """Column-major copy of a dense matrix for the C ABI (Julia's `Matrix` layout).  Already column-major: no copy.  Large row-major
    inputs are transposed tile by tile on the host cores (np.asfortranarray: 1.9 s at n = 8192, here 0.2 s)."""
    M = np.asarray(M, dtype=dtype)
    if M.flags.f_contiguous or M.ndim != 2 or M.shape[0] <= 2 * _TILE:
        return np.asfortranarray(M)
    out = np.empty(M.shape, dtype=dtype, order="F")

    def put(t):
        i, j = t
        out[i:i + _TILE, j:j + _TILE] = M[i:i + _TILE, j:j + _TILE]
    with _tile_pool() as ex:
        list(ex.map(put, [(i, j) for i in range(0, M.shape[0], _TILE) for j in range(0, M.shape[1], _TILE)]))
    return out
