"""ctypes binding of libfeastcuda (include/feastcuda.h).  No CPU fallback: a missing library or a
missing GPU raises immediately."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("FEASTCUDA_LIB", _HERE.parent / "lib" / "libfeastcuda.so"))

OK, ERR_ARG, ERR_CUDA, ERR_NCCL, ERR_UNSUPPORTED, ERR_STATE = range(6)
A, B = 0, 1
CSR, CSC = 0, 1
SYM, HERM, GEN = 0, 1, 2
SOLVER_DIRECT, SOLVER_BICGSTAB, SOLVER_MSLANCZOS = 0, 1, 2
KERN_NAMES = ("spmm_z", "lz_p1", "lz_upd", "lz_p2", "lz_res", "lz_cheb", "band_lu", "band_solve")
FILTER_REFERENCE, FILTER_TRUE = 0, 1
SHARD_NODES, SHARD_COLUMNS, SHARD_BALANCED = 0, 1, 2      # "rows" is a mode of the handle (feastcuda_set_row_sharding)


class SolverOpts(C.Structure):
    _fields_ = [("solver", C.c_int32), ("tol", C.c_double), ("maxiter", C.c_int32), ("restart", C.c_int32),
                ("inner_rel", C.c_double), ("ritz_guess", C.c_int32), ("filter", C.c_int32), ("shard", C.c_int32),
                ("check_every", C.c_int32), ("q0_real", C.c_int32), ("x_real", C.c_int32), ("inner_rel0", C.c_double),
                ("maxiter0", C.c_int32), ("keep_going", C.c_int32), ("adaptive", C.c_int32), ("mixed", C.c_int32), ("eps_floor", C.c_double),
                ("b_delta", C.c_double)]


class Stats(C.Structure):
    _fields_ = [("loops", C.c_int64), ("node_solves", C.c_int64), ("krylov_iters", C.c_int64), ("col_iters", C.c_int64),
                ("spmm_launches", C.c_int64), ("kernel_launches", C.c_int64), ("ortho_passes", C.c_int64),
                ("jacobi_sweeps", C.c_int64), ("allreduce_bytes", C.c_int64),
                ("ms_total", C.c_double), ("ms_h2d", C.c_double), ("ms_d2h", C.c_double), ("ms_solve", C.c_double),
                ("ms_ortho", C.c_double), ("ms_project", C.c_double), ("ms_eig", C.c_double), ("ms_resid", C.c_double),
                ("ms_allreduce", C.c_double), ("ms_spmm_sampled", C.c_double), ("spmm_sampled", C.c_int64),
                ("bytes_spmm_alg", C.c_double), ("node_iters", C.c_int64 * 128),
                ("lz_steps_p1", C.c_int64), ("lz_steps_p2", C.c_int64), ("ms_lz_p1", C.c_double), ("ms_lz_p2", C.c_double),
                ("ms_kern", C.c_double * 8), ("n_kern", C.c_int64 * 8), ("bytes_kern", C.c_double * 8), ("ms_dev_run", C.c_double),
                ("lz_steps_fp32", C.c_int64), ("cheb_degree", C.c_int64)]

    def as_dict(self):
        arrays = ("node_iters", "ms_kern", "n_kern", "bytes_kern")
        d = {k: getattr(self, k) for k, _ in self._fields_ if k not in arrays}
        for k in arrays:
            d[k] = list(getattr(self, k))
        return d


# feastcuda_apply_fn: void (*)(void* ctx, int64 n, int64 ncols, const double* X, int64 ldx, double* Y, int64 ldy, void* stream)
APPLY_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p)


class FeastCudaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libfeastcuda status {code}: {msg}")
        self.code = code


_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)
_vp = C.c_void_p

# name -> argtypes; every function returns int unless listed in _RESTYPES
SIGNATURES = {
    "feastcuda_create": [C.POINTER(_vp), C.c_int],
    "feastcuda_destroy": [_vp],
    "feastcuda_last_error": [_vp],
    "feastcuda_version": [],
    "feastcuda_feastinit": [_ip],
    "feastcuda_feastdefault": [_ip],
    "feastcuda_contour": [C.c_double, C.c_double, _ip, _dp, _dp],
    "feastcuda_gcontour": [C.c_double, C.c_double, C.c_double, _ip, _dp, _dp],
    "feastcuda_set_csr_d": [_vp, C.c_int, C.c_int64, C.c_int64, _ip, _ip, _dp, C.c_int, C.c_int, C.c_int],
    "feastcuda_set_csr_z": [_vp, C.c_int, C.c_int64, C.c_int64, _ip, _ip, _dp, C.c_int, C.c_int, C.c_int],
    "feastcuda_clear_b": [_vp],
    "feastcuda_set_matfree_d": [_vp, C.c_int64, C.c_void_p, C.c_void_p],
    "feastcuda_set_dense_d": [_vp, C.c_int, C.c_int64, _dp, C.c_int64, C.c_int],
    "feastcuda_set_dense_z": [_vp, C.c_int, C.c_int64, _dp, C.c_int64, C.c_int],
    "feastcuda_set_band_d": [_vp, C.c_int, C.c_int64, C.c_int64, _dp, C.c_int64, C.c_int],
    "feastcuda_set_band_z": [_vp, C.c_int, C.c_int64, C.c_int64, _dp, C.c_int64, C.c_int],
    "feastcuda_solve_interval": [_vp, C.c_double, C.c_double, C.c_int64, _ip, _dp, _dp, C.c_int64, _dp,
                                 C.POINTER(SolverOpts), _dp, _dp, _dp, _ip, _ip, _dp, _ip],
    "feastcuda_upload_subspace": [_vp, C.c_int64, _dp, C.c_int],
    "feastcuda_run_interval": [_vp, C.c_double, C.c_double, C.c_int64, _ip, _dp, _dp, C.c_int64, C.POINTER(SolverOpts),
                               _ip, _ip, _dp, _ip],
    "feastcuda_fetch_results": [_vp, C.c_int64, C.c_int, _dp, _dp, _dp],
    "feastcuda_solve_contour": [_vp, C.c_double, C.c_double, C.c_double, C.c_int64, _ip, _dp, _dp, C.c_int64, _dp,
                                C.POINTER(SolverOpts), _dp, _dp, _dp, _ip, _ip, _dp, _ip],
    "feastcuda_spmm_shifted": [_vp, C.c_double, C.c_double, C.c_int64, _dp, _dp],
    "feastcuda_apply": [_vp, C.c_int, C.c_int64, _dp, _dp],
    "feastcuda_block_solve": [_vp, C.c_double, C.c_double, C.c_int64, _dp, _dp, C.POINTER(SolverOpts), _dp, _ip, _dp],
    "feastcuda_accumulate": [_vp, C.c_double, C.c_double, C.c_int64, _dp, _dp],
    "feastcuda_orthonormalize": [_vp, C.c_int64, C.c_int64, _dp, C.c_double, _dp, _ip],
    "feastcuda_gram": [_vp, C.c_int64, C.c_int64, _dp, _dp, _dp],
    "feastcuda_reduced_eig": [_vp, C.c_int64, _dp, _dp, _dp, _dp, _ip],
    "feastcuda_residuals": [_vp, C.c_int64, _dp, _dp, _dp],
    "feastcuda_rowtransform": [_vp, C.c_int64, C.c_int64, C.c_int64, _dp, _dp, _dp],
    "feastcuda_eig_general": [_vp, C.c_int64, _dp, _dp, _dp, _dp],
    "feastcuda_nccl_unique_id": [C.c_char_p],
    "feastcuda_nccl_init": [_vp, C.c_int, C.c_int, C.c_char_p],
    "feastcuda_node_partition": [C.c_int64, C.c_int, C.c_int, _ip, _ip],
    "feastcuda_set_row_sharding": [_vp, C.c_int],
    "feastcuda_row_range": [_vp, _ip, _ip, _ip],
    "feastcuda_get_stats": [_vp, C.POINTER(Stats)],
    "feastcuda_reset_stats": [_vp],
}
_RESTYPES = {"feastcuda_last_error": C.c_char_p}


def load():
    """Load libfeastcuda.so (built in-tree by __graft_entry__.build()).  Fails loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise FeastCudaError(-1, f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                                 "(there is no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH), mode=C.RTLD_GLOBAL)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = _RESTYPES.get(name, C.c_int)
    _lib = lib
    return lib


def dptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


def iptr(a):
    return None if a is None else a.ctypes.data_as(_ip)


def check(rc, handle=None):
    if rc != OK:
        msg = load().feastcuda_last_error(handle)
        raise FeastCudaError(rc, msg.decode() if msg else "")


def fpm_array(fpm):
    """Reference fpm (list/array of 64 Ints) -> int64 buffer handed to the library."""
    a = np.ascontiguousarray(np.asarray(fpm, dtype=np.int64))
    if a.size < 64:
        raise ValueError("fpm array must have at least 64 elements")
    return a
