"""feastcuda -- host-side mirror of FeastKit.jl's API for the FEAST contour-integration path.

Julia is absent from the build image, so the host code a FeastKit.jl maintainer would write in Julia
(feastkit.jl_b200/julia/FeastCUDA.jl) is mirrored here in Python with the SAME names, argument order,
argument meaning and error behaviour (Julia's ``ArgumentError`` -> ``ValueError``; mutating ``name!`` ->
``name``).  Every function that computes goes through the C ABI of libfeastcuda (include/feastcuda.h);
nothing here falls back to the CPU, and nothing imports ``oracle/``.

file:line citations are relative to /root/reference/src.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import os

import numpy as np

from . import _lib as L
from ._host import column_major
from ._lib import (A, B, CSC, CSR, FILTER_REFERENCE, FILTER_TRUE, GEN, HERM, SHARD_BALANCED, SHARD_COLUMNS, SHARD_NODES,
                   SOLVER_BICGSTAB, SOLVER_DIRECT, SOLVER_MSLANCZOS, SYM, FeastCudaError, SolverOpts, Stats)

FEAST_UNINITIALIZED = -111


@dataclass
class FeastResult:
    """FeastResult{T,VT} / FeastGeneralResult{T} (core/feast_types.jl:85-108); arrays trimmed to M."""
    lambda_: np.ndarray
    q: np.ndarray
    M: int
    res: np.ndarray
    info: int
    epsout: float
    loop: int
    stats: dict = field(default_factory=dict)


FeastGeneralResult = FeastResult     # core/feast_types.jl:100-118: same fields, complex lambda


class _DevBlock:
    """A row-major (rows, cols) float64 block at a raw device address, exposed through __cuda_array_interface__."""

    def __init__(self, ptr, rows, cols, ld):
        self.__cuda_array_interface__ = {"shape": (int(rows), int(cols)), "typestr": "<f8", "data": (int(ptr), False), "version": 3,
                                         "strides": (int(ld) * 8, 8)}


def _device_view(ptr, rows, cols, ld, dev):
    import torch
    return torch.as_tensor(_DevBlock(ptr, rows, cols, ld), device=dev)


# ---- parameters / contours (core/feast_parameters.jl, core/feast_tools.jl) ---------------------------
def feastinit():
    """feastinit() -> 64 x -111 (core/feast_parameters.jl:20-24)."""
    fpm = np.zeros(64, dtype=np.int64)
    L.check(L.load().feastcuda_feastinit(L.iptr(fpm)))
    return [int(v) for v in fpm]


def feastinit_(fpm):
    """feastinit!(fpm) (core/feast_parameters.jl:7-18)."""
    if len(fpm) < 64:
        raise ValueError("fpm array must have at least 64 elements")
    for i in range(64):
        fpm[i] = FEAST_UNINITIALIZED
    return fpm


def feastdefault_(fpm):
    """feastdefault!(fpm) (core/feast_parameters.jl:41-386); invalid entries raise ValueError."""
    a = L.fpm_array(fpm)
    if L.load().feastcuda_feastdefault(L.iptr(a)) != L.OK:
        raise ValueError("Invalid fpm parameter (feastdefault!)")
    for i in range(64):
        fpm[i] = int(a[i])
    return fpm


def feast_tolerance(fpm, dtype=np.float64):
    """core/feast_parameters.jl:391-405"""
    e = fpm[2]
    tol = 1e-12 if (e < 0 or e > 16) else 10.0 ** (-e)
    if np.dtype(dtype) == np.float32:
        return max(float(np.float32(tol)), float(np.sqrt(np.finfo(np.float32).eps)))
    return tol


def feast_contour(Emin, Emax, fpm):
    """feast_contour(Emin, Emax, fpm) -> (Zne, Wne) (core/feast_tools.jl:212-284)."""
    a = L.fpm_array(fpm)
    if a[1] == FEAST_UNINITIALIZED or a[1] <= 0:
        feastdefault_(fpm)
        a = L.fpm_array(fpm)
    ne = int(a[1])
    Z = np.zeros(ne, dtype=np.complex128)
    W = np.zeros(ne, dtype=np.complex128)
    rc = L.load().feastcuda_contour(float(Emin), float(Emax), L.iptr(a), Z.ctypes.data_as(L._dp), W.ctypes.data_as(L._dp))
    L.check(rc)
    return Z, W


def feast_gcontour(Emid, r, fpm):
    """feast_gcontour(Emid, r, fpm) -> (Zne, Wne) (core/feast_tools.jl:286-371)."""
    a = L.fpm_array(fpm)
    if a[7] == FEAST_UNINITIALIZED or a[7] <= 0:
        feastdefault_(fpm)
        a = L.fpm_array(fpm)
    ne = int(a[7])
    Z = np.zeros(ne, dtype=np.complex128)
    W = np.zeros(ne, dtype=np.complex128)
    Emid = complex(Emid)
    L.check(L.load().feastcuda_gcontour(Emid.real, Emid.imag, float(r), L.iptr(a), Z.ctypes.data_as(L._dp), W.ctypes.data_as(L._dp)))
    return Z, W


def check_feast_srci_input(N, M0, Emin, Emax, fpm):
    """core/feast_aux.jl:369-399 -- thrown host-side before the ccall, as in the reference."""
    if N <= 0:
        raise ValueError("Matrix size N must be positive")
    if M0 <= 0 or M0 > N:
        raise ValueError("Number of eigenvalues M0 must be between 1 and N")
    if Emin >= Emax:
        raise ValueError("Search interval [Emin, Emax] must be valid")
    if len(fpm) < 64:
        raise ValueError("fpm array must have at least 64 elements")
    if 0 < fpm[1] < 3:
        raise ValueError("Number of integration points must be at least 3")
    return True


def feast_inside_contour(lam, Emin, Emax):
    """feast_inside_contour (core/feast_tools.jl:619-621): closed interval."""
    return Emin <= lam <= Emax


def feast_inside_gcontour(lam, Emid, r, fpm=None):
    """feast_inside_gcontour(lambda, Emid, r; fpm) (core/feast_tools.jl:623-650): membership in the ellipse of half-axes r and
    r*fpm[18]/100 rotated by fpm[19] degrees about Emid -- the same rule as host_inside_gcontour in csrc/host_math.hpp."""
    import math
    w = complex(lam) - complex(Emid)
    aspect, rot = 1.0, 0.0
    if fpm is not None and len(fpm) >= 19:
        if fpm[17] > 0:
            aspect = fpm[17] * 0.01
        if fpm[18] != 0:
            rot = (fpm[18] / 180.0) * math.pi
    if rot != 0.0:
        w *= complex(math.cos(-rot), math.sin(-rot))
    x, y = w.real / r, w.imag / (r * aspect)
    return x * x + y * y <= 1.0


def check_feast_grci_input(N, M0, Emid, r, fpm):
    """core/feast_aux.jl:401-425"""
    if N <= 0:
        raise ValueError("Matrix size N must be positive")
    if M0 <= 0 or M0 > N:
        raise ValueError("Number of eigenvalues M0 must be between 1 and N")
    if r <= 0:
        raise ValueError("Search radius r must be positive")
    if len(fpm) < 64:
        raise ValueError("fpm array must have at least 64 elements")
    return True


def node_partition(ne, nranks, rank):
    """Block node distribution (parallel/feast_mpi.jl:36-43); returns (start, count), 0-based."""
    s = np.zeros(1, dtype=np.int64)
    c = np.zeros(1, dtype=np.int64)
    L.check(L.load().feastcuda_node_partition(int(ne), int(nranks), int(rank), L.iptr(s), L.iptr(c)))
    return int(s[0]), int(c[0])


# ---- engine handle --------------------------------------------------------------------------------------
def _as_z(a):
    return np.ascontiguousarray(a, dtype=np.complex128)


def _colmajor_z(X):
    return np.asfortranarray(np.asarray(X, dtype=np.complex128))


def _view_d(a):
    return a.view(np.float64) if a.dtype == np.complex128 else a


class Engine:
    """One libfeastcuda handle bound to one GPU."""

    def __init__(self, device=0):
        self.lib = L.load()
        self.h = L._vp()
        L.check(self.lib.feastcuda_create(C.byref(self.h), int(device)))
        self.device = device
        self.n = 0
        self.distributed = False

    def close(self):
        if self.h:
            self.lib.feastcuda_destroy(self.h)
            self.h = L._vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        err = getattr(self, "_matfree_error", None)
        if err is not None:            # an exception inside a Python matrix-free callback cannot cross the C frames: re-raise it here
            self._matfree_error = None
            raise err
        L.check(rc, self.h)

    # -- operators
    def set_sparse(self, which, M, structure, fmt=None):
        """SparseMatrixCSC analogue: scipy CSC (Julia's layout) or CSR."""
        import scipy.sparse as sp
        if fmt is None:
            fmt = CSR if sp.isspmatrix_csr(M) else CSC
        M = M.tocsr() if fmt == CSR else M.tocsc()
        M.sort_indices()
        n = M.shape[0]
        if M.shape[0] != M.shape[1]:
            raise ValueError("Matrix must be square")
        ptr = np.ascontiguousarray(M.indptr, dtype=np.int64)
        idx = np.ascontiguousarray(M.indices, dtype=np.int64)
        if np.iscomplexobj(M.data):
            val = _as_z(M.data)
            self._ck(self.lib.feastcuda_set_csr_z(self.h, which, n, M.nnz, L.iptr(ptr), L.iptr(idx), val.ctypes.data_as(L._dp), 0, fmt, structure))
        else:
            val = np.ascontiguousarray(M.data, dtype=np.float64)
            self._ck(self.lib.feastcuda_set_csr_d(self.h, which, n, M.nnz, L.iptr(ptr), L.iptr(idx), L.dptr(val), 0, fmt, structure))
        if which == A:
            self.n = n

    def set_dense(self, which, M, structure):
        M = np.asarray(M)
        n = M.shape[0]
        if M.ndim != 2 or M.shape[1] != n:
            raise ValueError("Matrix must be square")
        if np.iscomplexobj(M):
            a = column_major(M, np.complex128)
            self._ck(self.lib.feastcuda_set_dense_z(self.h, which, n, a.ctypes.data_as(L._dp), n, structure))
        else:
            # a row-major real matrix declared SYMMETRIC is its own column-major image: no host copy at all
            a = np.ascontiguousarray(M, dtype=np.float64) if (structure == SYM and M.flags.c_contiguous) else column_major(M, np.float64)
            self._ck(self.lib.feastcuda_set_dense_d(self.h, which, n, L.dptr(a), n, structure))
        if which == A:
            self.n = n

    def set_band(self, which, AB, k, structure):
        AB = np.asarray(AB)
        n = AB.shape[1]
        ldab = AB.shape[0]
        if np.iscomplexobj(AB):
            a = np.asfortranarray(AB, dtype=np.complex128)
            self._ck(self.lib.feastcuda_set_band_z(self.h, which, n, int(k), a.ctypes.data_as(L._dp), ldab, structure))
        else:
            a = np.asfortranarray(AB, dtype=np.float64)
            self._ck(self.lib.feastcuda_set_band_d(self.h, which, n, int(k), L.dptr(a), ldab, structure))
        if which == A:
            self.n = n

    def clear_b(self):
        self._ck(self.lib.feastcuda_clear_b(self.h))

    def set_matfree(self, n, apply, ctx=None):
        """Matrix-free real symmetric A (B = I): `apply` is either a C function pointer of type feastcuda_apply_fn (an int
        address or a ctypes function, with `ctx` passed through as its void* context -- e.g. the example operator in
        examples/matfree_laplacian.cu) or a Python callable apply(Y, X) that receives the device blocks as torch CUDA tensor
        views of shape (n, ncols) and must fill Y = A @ X with torch operations (they are enqueued on the library's stream)."""
        if callable(apply) and not isinstance(apply, C._CFuncPtr):
            user = apply

            def trampoline(_ctx, nn, ncols, xp, ldx, yp, ldy, stream):
                import torch
                dev = torch.device("cuda", self.device)
                if getattr(self, "_matfree_error", None) is not None:
                    return
                try:
                    with torch.cuda.stream(torch.cuda.ExternalStream(stream or 0, device=dev)):
                        X = _device_view(xp, nn, ncols, ldx, dev)
                        Y = _device_view(yp, nn, ncols, ldy, dev)
                        user(Y, X)
                except BaseException as exc:   # noqa: BLE001
                    self._matfree_error = exc

            fn = L.APPLY_FN(trampoline)
            self._matfree_keep = (fn, user)
            addr, cx = C.cast(fn, C.c_void_p), None
        else:
            self._matfree_keep = (apply, ctx)
            addr = C.cast(apply, C.c_void_p) if isinstance(apply, C._CFuncPtr) else C.c_void_p(int(apply))
            cx = ctx if (ctx is None or isinstance(ctx, (int, C.c_void_p))) else C.cast(C.pointer(ctx), C.c_void_p)
        self._ck(self.lib.feastcuda_set_matfree_d(self.h, int(n), addr, cx))
        self.n = int(n)

    # -- multi-GPU: one process per GPU, NCCL id distributed through torch.distributed
    def init_distributed(self):
        import torch.distributed as dist
        if not dist.is_initialized() or dist.get_world_size() == 1 or self.distributed:
            return
        import torch
        buf = C.create_string_buffer(128)
        if dist.get_rank() == 0:
            L.check(self.lib.feastcuda_nccl_unique_id(buf))
        t = torch.tensor(list(buf.raw), dtype=torch.uint8)
        if dist.get_backend() == "nccl":
            t = t.cuda(self.device)
        dist.broadcast(t, 0)
        raw = bytes(t.cpu().tolist())
        self._ck(self.lib.feastcuda_nccl_init(self.h, dist.get_world_size(), dist.get_rank(), raw))
        self.distributed = True

    def set_row_sharding(self, on=True):
        """Row-sharded multi-GPU mode (feastcuda_set_row_sharding): call after init_distributed; a no-op on one rank."""
        self._ck(self.lib.feastcuda_set_row_sharding(self.h, int(bool(on))))
        self.row_sharded = bool(on) and (self.distributed or bool(os.environ.get("FEASTCUDA_FORCE_ROWS")))

    def row_range(self):
        """(row0, nrows, nglobal): this rank's block of rows (the whole matrix without row sharding)."""
        a = np.zeros(3, dtype=np.int64)
        self._ck(self.lib.feastcuda_row_range(self.h, L.iptr(a[0:1]), L.iptr(a[1:2]), L.iptr(a[2:3])))
        return int(a[0]), int(a[1]), int(a[2])

    def _gather_rows(self, X):
        """Row-sharded results: every rank holds its own rows of X; fill in the others (torch.distributed all_reduce of the zero-padded
        block -- the drop-in API returns the full eigenvectors on every rank like the reference's MPI drivers)."""
        import torch
        import torch.distributed as dist
        row0, nrows, _ = self.row_range()
        full = np.zeros_like(X)
        full[row0:row0 + nrows] = X[row0:row0 + nrows]
        t = torch.from_numpy(np.ascontiguousarray(full).view(np.float64))
        if dist.get_backend() == "nccl":
            t = t.cuda(self.device)
        dist.all_reduce(t)
        return t.cpu().numpy().view(X.dtype).reshape(X.shape)

    # -- solves
    @staticmethod
    def make_opts(solver="bicgstab", solver_tol=0.0, solver_maxiter=500, solver_restart=3, inner_rel=0.0, ritz_guess=False,
                  filter="reference", shard="nodes", check_every=8, q0_real=False, x_real=False, inner_rel0=0.0, maxiter0=0,
                  keep_going=False, adaptive=False, eps_floor=0.0, mixed=False, b_delta=0.0):
        o = SolverOpts()
        o.solver = {"direct": SOLVER_DIRECT, "bicgstab": SOLVER_BICGSTAB, "mslanczos": SOLVER_MSLANCZOS}[solver]
        o.tol = float(solver_tol)
        o.maxiter = int(solver_maxiter)
        o.restart = int(solver_restart)
        o.inner_rel = float(inner_rel)
        o.ritz_guess = int(bool(ritz_guess))
        o.filter = FILTER_TRUE if filter == "true" else FILTER_REFERENCE
        o.shard = {"nodes": SHARD_NODES, "columns": SHARD_COLUMNS, "balanced": SHARD_BALANCED, "rows": SHARD_COLUMNS}[shard]
        o.check_every = int(check_every)
        o.q0_real = int(bool(q0_real))
        o.x_real = int(bool(x_real))
        o.inner_rel0 = float(inner_rel0)
        o.maxiter0 = int(maxiter0)
        o.keep_going = int(bool(keep_going))
        o.adaptive = int(bool(adaptive))
        o.eps_floor = float(eps_floor)
        o.b_delta = float(b_delta)
        o.mixed = 2 if mixed == "fpm" else int(bool(mixed))      # "fpm": follow fpm[42] (core/feast_parameters.jl:316-319)
        return o

    def upload_subspace(self, M0, Q0=None):
        if Q0 is None:
            self._ck(self.lib.feastcuda_upload_subspace(self.h, int(M0), None, 0))
            return
        Q0 = np.asarray(Q0)
        if Q0.shape != (self.n, M0):
            raise ValueError("Q0 must be N x M0")
        if np.iscomplexobj(Q0):
            q = _colmajor_z(Q0)
            self._ck(self.lib.feastcuda_upload_subspace(self.h, int(M0), q.ctypes.data_as(L._dp), 0))
        else:
            q = np.asfortranarray(Q0, dtype=np.float64)
            self._ck(self.lib.feastcuda_upload_subspace(self.h, int(M0), L.dptr(q), 1))

    def run_interval(self, Emin, Emax, M0, fpm, Zne, Wne, opts):
        a = L.fpm_array(fpm)
        Z = _as_z(Zne)
        W = _as_z(Wne)
        M = np.zeros(1, dtype=np.int64)
        info = np.zeros(1, dtype=np.int64)
        loop = np.zeros(1, dtype=np.int64)
        eps = np.zeros(1, dtype=np.float64)
        self._ck(self.lib.feastcuda_run_interval(self.h, float(Emin), float(Emax), int(M0), L.iptr(a), Z.ctypes.data_as(L._dp),
                                                 W.ctypes.data_as(L._dp), len(Z), C.byref(opts), L.iptr(M), L.iptr(info), L.dptr(eps), L.iptr(loop)))
        for i in range(64):
            fpm[i] = int(a[i])
        return int(M[0]), int(info[0]), float(eps[0]), int(loop[0])

    def fetch_results(self, M0, M, x_real):
        lam = np.zeros(M0, dtype=np.float64)
        res = np.zeros(M0, dtype=np.float64)
        X = np.zeros((self.n, max(M, 1)), dtype=np.float64 if x_real else np.complex128, order="F")
        self._ck(self.lib.feastcuda_fetch_results(self.h, int(M0), int(x_real), L.dptr(lam), X.ctypes.data_as(L._dp), L.dptr(res)))
        return lam[:M].copy(), X[:, :M].copy(), res[:M].copy()

    def solve_interval(self, Emin, Emax, M0, fpm, Zne, Wne, Q0=None, x_real=False, **kw):
        """One C-ABI call with HOST buffers in and out (the end-to-end path).  The returned stats describe this solve.
        shard="rows" (multi-rank runs): row-sharded mode; gather_rows=False leaves only this rank's rows of q filled."""
        gather_rows = kw.pop("gather_rows", True)
        reuse_output = kw.pop("reuse_output", False)    # True: q is a view of a buffer the NEXT solve overwrites (no 8 n M0-byte allocation per call)
        if self.distributed:
            self.set_row_sharding(kw.get("shard") == "rows")
        self.reset_stats()
        a = L.fpm_array(fpm)
        Z = _as_z(Zne)
        W = _as_z(Wne)
        q0_real = Q0 is not None and not np.iscomplexobj(Q0)
        opts = self.make_opts(q0_real=q0_real, x_real=x_real, **kw)
        q = None
        if Q0 is not None:
            Q0 = np.asarray(Q0)
            if Q0.shape != (self.n, M0):
                raise ValueError("Q0 must be N x M0")
            q = np.asfortranarray(Q0, dtype=np.float64) if q0_real else _colmajor_z(Q0)
        lam = np.zeros(M0, dtype=np.float64)
        res = np.zeros(M0, dtype=np.float64)
        xkey = (self.n, int(M0), bool(x_real))
        if reuse_output and getattr(self, "_xbuf_key", None) == xkey:
            X = self._xbuf
        else:
            X = np.zeros((self.n, M0), dtype=np.float64 if x_real else np.complex128, order="F")
            if reuse_output:
                self._xbuf, self._xbuf_key = X, xkey
        M = np.zeros(1, dtype=np.int64)
        info = np.zeros(1, dtype=np.int64)
        loop = np.zeros(1, dtype=np.int64)
        eps = np.zeros(1, dtype=np.float64)
        self._ck(self.lib.feastcuda_solve_interval(self.h, float(Emin), float(Emax), int(M0), L.iptr(a), Z.ctypes.data_as(L._dp),
                                                   W.ctypes.data_as(L._dp), len(Z), None if q is None else q.ctypes.data_as(L._dp),
                                                   C.byref(opts), L.dptr(lam), X.ctypes.data_as(L._dp), L.dptr(res), L.iptr(M),
                                                   L.iptr(info), L.dptr(eps), L.iptr(loop)))
        for i in range(64):
            fpm[i] = int(a[i])
        m = int(M[0])
        Xm = X[:, :m] if reuse_output else X[:, :m].copy()
        if getattr(self, "row_sharded", False) and gather_rows and m > 0:
            Xm = self._gather_rows(Xm)
        return FeastResult(lam[:m].copy(), Xm, m, res[:m].copy(), int(info[0]), float(eps[0]), int(loop[0]), self.stats())

    def solve_contour(self, Emid, r, M0, fpm, Zne, Wne, Q0=None, **kw):
        """General (non-Hermitian) solve, one C-ABI call with host buffers (feastcuda_solve_contour)."""
        self.reset_stats()
        a = L.fpm_array(fpm)
        Z = _as_z(Zne)
        W = _as_z(Wne)
        q0_real = Q0 is not None and not np.iscomplexobj(Q0)
        opts = self.make_opts(q0_real=q0_real, x_real=False, **kw)
        q = None
        if Q0 is not None:
            Q0 = np.asarray(Q0)
            if Q0.shape != (self.n, M0):
                raise ValueError("Q0 must be N x M0")
            q = np.asfortranarray(Q0, dtype=np.float64) if q0_real else _colmajor_z(Q0)
        lam = np.zeros(M0, dtype=np.complex128)
        res = np.zeros(M0, dtype=np.float64)
        X = np.zeros((self.n, M0), dtype=np.complex128, order="F")
        M = np.zeros(1, dtype=np.int64)
        info = np.zeros(1, dtype=np.int64)
        loop = np.zeros(1, dtype=np.int64)
        eps = np.zeros(1, dtype=np.float64)
        Emid = complex(Emid)
        self._ck(self.lib.feastcuda_solve_contour(self.h, Emid.real, Emid.imag, float(r), int(M0), L.iptr(a), Z.ctypes.data_as(L._dp),
                                                  W.ctypes.data_as(L._dp), len(Z), None if q is None else q.ctypes.data_as(L._dp),
                                                  C.byref(opts), lam.ctypes.data_as(L._dp), X.ctypes.data_as(L._dp), L.dptr(res),
                                                  L.iptr(M), L.iptr(info), L.dptr(eps), L.iptr(loop)))
        for i in range(64):
            fpm[i] = int(a[i])
        m = int(M[0])
        return FeastResult(lam[:m].copy(), X[:, :m].copy(), m, res[:m].copy(), int(info[0]), float(eps[0]), int(loop[0]), self.stats())

    # -- stage-level entry points (host buffers)
    def spmm_shifted(self, z, X):
        X = _colmajor_z(X)
        Y = np.zeros_like(X, order="F")
        z = complex(z)
        self._ck(self.lib.feastcuda_spmm_shifted(self.h, z.real, z.imag, X.shape[1], X.ctypes.data_as(L._dp), Y.ctypes.data_as(L._dp)))
        return Y

    def apply(self, which, X):
        X = _colmajor_z(X)
        Y = np.zeros_like(X, order="F")
        self._ck(self.lib.feastcuda_apply(self.h, which, X.shape[1], X.ctypes.data_as(L._dp), Y.ctypes.data_as(L._dp)))
        return Y

    def block_solve(self, z, RHS, X0=None, **kw):
        RHS = _colmajor_z(RHS)
        m = RHS.shape[1]
        X = np.zeros_like(RHS, order="F")
        x0 = None if X0 is None else _colmajor_z(X0)
        iters = np.zeros(m, dtype=np.int64)
        resid = np.zeros(m, dtype=np.float64)
        opts = self.make_opts(**kw)
        z = complex(z)
        self._ck(self.lib.feastcuda_block_solve(self.h, z.real, z.imag, m, RHS.ctypes.data_as(L._dp),
                                                None if x0 is None else x0.ctypes.data_as(L._dp), C.byref(opts),
                                                X.ctypes.data_as(L._dp), L.iptr(iters), L.dptr(resid)))
        return X, iters, resid

    def accumulate(self, w, Y, Qacc):
        Y = _colmajor_z(Y)
        Q = _colmajor_z(Qacc).copy(order="F")
        w = complex(w)
        self._ck(self.lib.feastcuda_accumulate(self.h, w.real, w.imag, Y.shape[1], Y.ctypes.data_as(L._dp), Q.ctypes.data_as(L._dp)))
        return Q

    def orthonormalize(self, W, rank_tol=0.0):
        W = _colmajor_z(W)
        n, m = W.shape
        Q = np.zeros_like(W, order="F")
        rank = np.zeros(1, dtype=np.int64)
        self._ck(self.lib.feastcuda_orthonormalize(self.h, n, m, W.ctypes.data_as(L._dp), float(rank_tol), Q.ctypes.data_as(L._dp), L.iptr(rank)))
        r = int(rank[0])
        return Q[:, :r].copy(), r

    def gram(self, X, Y):
        X = _colmajor_z(X)
        Y = _colmajor_z(Y)
        n, m = X.shape
        Cm = np.zeros((m, m), dtype=np.complex128, order="F")
        self._ck(self.lib.feastcuda_gram(self.h, n, m, X.ctypes.data_as(L._dp), Y.ctypes.data_as(L._dp), Cm.ctypes.data_as(L._dp)))
        return Cm

    def reduced_eig(self, Sq, Aq=None):
        Sq = _colmajor_z(Sq)
        r = Sq.shape[0]
        Aq_ = None if Aq is None else _colmajor_z(Aq)
        lam = np.zeros(r, dtype=np.float64)
        V = np.zeros((r, r), dtype=np.complex128, order="F")
        sw = np.zeros(1, dtype=np.int64)
        self._ck(self.lib.feastcuda_reduced_eig(self.h, r, Sq.ctypes.data_as(L._dp), None if Aq_ is None else Aq_.ctypes.data_as(L._dp),
                                                L.dptr(lam), V.ctypes.data_as(L._dp), L.iptr(sw)))
        return lam, V, int(sw[0])

    def rowtransform(self, X, T):
        """Y = X @ T on the device (back-projection of the RCI kernels / drivers)."""
        X = _colmajor_z(X)
        T = _colmajor_z(T)
        n, a = X.shape
        b = T.shape[1]
        Y = np.zeros((n, b), dtype=np.complex128, order="F")
        self._ck(self.lib.feastcuda_rowtransform(self.h, n, a, b, X.ctypes.data_as(L._dp), T.ctypes.data_as(L._dp), Y.ctypes.data_as(L._dp)))
        return Y

    def eig_general(self, A, B=None):
        """eigen(A, B) of a small general pencil: (lambda, V) with unit-norm columns."""
        A = _colmajor_z(A)
        r = A.shape[0]
        B_ = None if B is None else _colmajor_z(B)
        lam = np.zeros(r, dtype=np.complex128)
        V = np.zeros((r, r), dtype=np.complex128, order="F")
        self._ck(self.lib.feastcuda_eig_general(self.h, r, A.ctypes.data_as(L._dp), None if B_ is None else B_.ctypes.data_as(L._dp),
                                                lam.ctypes.data_as(L._dp), V.ctypes.data_as(L._dp)))
        return lam, V

    def residuals(self, X, lam):
        X = _colmajor_z(X)
        lam = _as_z(lam)
        res = np.zeros(X.shape[1], dtype=np.float64)
        self._ck(self.lib.feastcuda_residuals(self.h, X.shape[1], X.ctypes.data_as(L._dp), lam.ctypes.data_as(L._dp), L.dptr(res)))
        return res

    def stats(self):
        s = Stats()
        self._ck(self.lib.feastcuda_get_stats(self.h, C.byref(s)))
        return s.as_dict()

    def reset_stats(self):
        self._ck(self.lib.feastcuda_reset_stats(self.h))


_ENGINES = {}


def default_engine(device=None):
    """Process-wide handle for `device` (default: LOCAL_RANK under torchrun, else 0)."""
    import os
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if device not in _ENGINES:
        _ENGINES[device] = Engine(device)
    return _ENGINES[device]


from .api import *  # noqa: E402,F401,F403  (reference-named drivers)
from .families import *  # noqa: E402,F401,F403  (complex-symmetric and polynomial names)
from .utils import *  # noqa: E402,F401,F403  (host-side helpers of the API surface)
from .fixtures import *  # noqa: E402,F401,F403  (readers of the FEAST example-system files)
from .rci import (FeastRCIState, Ref, dfeast_srci, feast_grci, feast_grcix, feast_hrci, feast_hrcix, feast_srci, feast_srcix,  # noqa: E402,F401
                  ifeast_grci, ifeast_hrci, ifeast_srci, pdfeast_srci, zfeast_grci, zfeast_hrci)
