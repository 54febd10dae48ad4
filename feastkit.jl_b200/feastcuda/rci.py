"""Reverse-communication (RCI) state machines -- the host-side contract of kernel/feast_kernel.jl.

The caller owns the shifted factorisations/solves and the mat-vecs; these functions keep the reference's job codes,
argument order and state handling:

    ijob = -1 (INIT) -> 10 FACTORIZE (Ze*B - A) -> 11 SOLVE (workc <- solve(B*work)) -> ... every node ... ->
    30 MULT_A (work[:, :M] <- A*q[:, :M]) -> 0 DONE  or  10 again for the next refinement loop

`feast_srci` (real symmetric, kernel/feast_kernel.jl:7-293) and `feast_hrci` (complex Hermitian, :397-644) are the moment
(S-MOM / H-MOM) variants: per node Q_proj += 2 w_e Y, zAq += 2 w_e Q0^H Y, zSq += 2 w_e z_e Q0^H Y; after the sweep
eigen(Sq, Aq), q = Q_proj V, stable inside-first partition, residuals from the caller's A*q.  `feast_grci` (:646-962) is the
general one-sided variant on the full contour (q += w_e Y, MULT_B and MULT_A requests for the projections, eigen(Aq, Sq)).

Reference behaviour kept as is: `feast_hrci` accumulates the half-contour sums 2 w_e Y WITHOUT a Hermitian part
(kernel/feast_kernel.jl:516-524), so its moments are Q0^H g(A) Q0 with the complex g(x) = sum_e 2 w_e / (z_e - x) and
real(eigen(zSq, zAq)) equals x + Re(c / g(x)), c = sum_e 2 w_e -- exact only at the centre of the interval.  The reference has
no internal caller and tests only its INIT handshake; the mirror reproduces the same numbers (tests/test_gpu_solve.py checks
them against a NumPy restatement of those lines), it does not "fix" the routine.  When Aq is rank deficient (M0 larger than
the number of eigenvalues inside) LAPACK's QZ returns infinite/arbitrary values for the null directions; here they come back
as +inf and are partitioned outside.

Julia's `Ref`s are `Ref` objects here (`.v`); arrays are NumPy arrays mutated in place.  The n x M0 arithmetic of the
kernel (accumulation, moments, back-projection, residual norms) runs on the GPU through libfeastcuda's stage-level entry
points (`feastcuda_accumulate`, `feastcuda_gram`, `feastcuda_rowtransform`), the small reduced eigenproblem through
`feastcuda_eig_general`; nothing here computes with the CPU beyond O(M0^2) bookkeeping.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

RCI_INIT, RCI_DONE, RCI_FACTORIZE, RCI_SOLVE, RCI_MULT_A, RCI_MULT_B = -1, 0, 10, 11, 30, 40
SUCCESS, ERR_N, ERR_M0, ERR_EMIN_EMAX, ERR_EMID_R, ERR_NO_CONV, ERR_LAPACK = 0, 1, 2, 3, 4, 5, 8


class Ref:
    """Base.RefValue stand-in."""
    __slots__ = ("v",)

    def __init__(self, v=0):
        self.v = v

    def __repr__(self):
        return f"Ref({self.v!r})"


@dataclass
class FeastRCIState:
    """FeastSRCIState / FeastHRCIState / FeastGRCIState (core/feast_types.jl:120-192)."""
    Zne: np.ndarray = None
    Wne: np.ndarray = None
    ne: int = 0
    e: int = 1
    initialized: bool = False
    Q0: np.ndarray = None
    Q_proj: np.ndarray = None
    zAq: np.ndarray = None
    zSq: np.ndarray = None
    M: int = 0
    extra: dict = field(default_factory=dict)


def _engine(engine):
    from . import default_engine
    return engine if engine is not None else default_engine()


def _seed_subspace(N, M0, complex_storage):
    """_feast_seeded_subspace! (core/feast_tools.jl:6-43): deterministic real Gaussian columns of unit norm.  Julia's RNG
    stream cannot be reproduced here; the seed depends on (N, M0) like the reference's."""
    rng = np.random.default_rng(abs(hash((N, M0))) % (2 ** 32))
    Q = rng.standard_normal((N, M0))
    Q /= np.linalg.norm(Q, axis=0)
    return Q.astype(np.complex128) if complex_storage else Q


def _moment_rci(ijob, N, Ze, work, workc, Aq, Sq, fpm, epsout, loop, Emin, Emax, M0, lambda_, q, mode, res, info, state, hermitian,
                engine, contour=None):
    from . import feast_contour, feast_tolerance, feastdefault_
    # srci: trial block / A*q travel in the real `work`, solutions in `workc`; hrci: everything travels in the complex `workc`
    io = workc if hermitian else work
    if ijob.v == RCI_INIT:
        feastdefault_(fpm)
        info.v = SUCCESS
        if N <= 0:
            info.v = ERR_N
            return
        if M0 <= 0 or M0 > N:
            info.v = ERR_M0
            return
        if Emin >= Emax:
            info.v = ERR_EMIN_EMAX
            return
        Z, W = _custom(contour) if contour is not None else feast_contour(Emin, Emax, fpm)
        state.Zne, state.Wne, state.ne, state.e, state.initialized = Z.copy(), W.copy(), len(Z), 1, True
        fpm[49], fpm[50], fpm[51], fpm[52] = 1, len(Z), 0, 1          # fpm[50..53] (1-based): node counter, ne, M, initialised
        loop.v = 0
        user = io[:, :M0].copy() if fpm[4] == 1 else None             # fpm[5]: user-provided initial subspace
        for a in (Aq, Sq, lambda_, q, res, workc, work):
            a[...] = 0
        if user is not None:
            for j in range(M0):
                nrm = np.linalg.norm(user[:, j])
                io[:, j] = user[:, j] / nrm if nrm > 0 else _seed_subspace(N, 1, hermitian)[:, 0]
        else:
            io[:, :M0] = _seed_subspace(N, M0, hermitian)
        state.Q0 = io[:, :M0].astype(np.complex128 if hermitian else np.float64).copy()
        state.Q_proj = np.zeros((N, M0), dtype=np.complex128)
        state.zAq = np.zeros((M0, M0), dtype=np.complex128)
        state.zSq = np.zeros((M0, M0), dtype=np.complex128)
        Ze.v = complex(Z[0])
        ijob.v = RCI_FACTORIZE
        return
    if ijob.v == RCI_FACTORIZE:
        ijob.v = RCI_SOLVE
        io[:, :M0] = state.Q0
        return
    if ijob.v == RCI_SOLVE:
        eng = _engine(engine)
        e, ne = state.e, state.ne
        if e == 1:
            state.Q_proj[...] = 0
            state.zAq[...] = 0
            state.zSq[...] = 0
        weight = 2 * state.Wne[e - 1]
        Y = np.asarray(workc[:, :M0], dtype=np.complex128)
        state.Q_proj[:, :] = eng.accumulate(weight, Y, state.Q_proj)                 # Q_proj += 2 w_e Y        (GPU)
        moment = eng.gram(state.Q0.astype(np.complex128), Y)                         # Q0^H Y                   (GPU)
        state.zAq += weight * moment
        state.zSq += weight * state.Zne[e - 1] * moment
        fpm[49] = e + 1
        state.e = e + 1
        if e < ne:
            Ze.v = complex(state.Zne[e])
            ijob.v = RCI_FACTORIZE
            return
        fpm[49] = 1
        state.e = 1
        if hermitian:
            Aq[:M0, :M0], Sq[:M0, :M0] = state.zAq, state.zSq
            A_red, S_red = state.zAq, state.zSq
        else:
            Aq[:M0, :M0], Sq[:M0, :M0] = state.zAq.real, state.zSq.real     # imaginary parts cancel over the mirrored half contour
            A_red, S_red = state.zAq.real.astype(np.complex128), state.zSq.real.astype(np.complex128)
        try:
            lam_red, V = eng.eig_general(S_red, A_red)                               # eigen(Sq, Aq)
        except Exception as err:                                                      # kernel/feast_kernel.jl:284-289 catch
            state.extra["error"] = str(err)
            info.v = ERR_LAPACK
            ijob.v = RCI_DONE
            fpm[52] = 0
            state.initialized = False
            return
        lam_red = lam_red.real
        if not hermitian:
            # real pencil: eigenvectors of real eigenvalues are real up to a phase; rotate each to its real representative
            for k in range(M0):
                j = int(np.argmax(np.abs(V[:, k])))
                V[:, k] *= np.conj(V[j, k]) / abs(V[j, k])
            V = V.real.astype(np.complex128)
            Qp = state.Q_proj.real.astype(np.complex128)
        else:
            Qp = state.Q_proj
        X = eng.rowtransform(Qp, V)                                                   # q = Q_proj V             (GPU)
        inside = [i for i in range(M0) if Emin <= lam_red[i] <= Emax]
        perm = inside + [i for i in range(M0) if not (Emin <= lam_red[i] <= Emax)]
        M = len(inside)
        lambda_[:M0] = lam_red[perm]
        q[:, :M0] = X[:, perm] if hermitian else X[:, perm].real
        fpm[51] = M
        state.M = M
        if M == 0:
            info.v = ERR_NO_CONV
            ijob.v = RCI_DONE
            fpm[52] = 0
            state.initialized = False
            return
        ijob.v = RCI_MULT_A
        mode.v = M
        return
    if ijob.v == RCI_MULT_A:
        eng = _engine(engine)
        M = fpm[51]
        # ||work_j - lambda_j q_j|| / max(|lambda_j|, 1)  (B is not part of the RCI residual, kernel/feast_kernel.jl:250)
        Lq = eng.rowtransform(np.asarray(q[:, :M], dtype=np.complex128), np.diag(lambda_[:M]).astype(np.complex128))
        R = eng.accumulate(-1.0, Lq, np.asarray(io[:, :M], dtype=np.complex128))
        nrm2 = np.real(np.diag(eng.gram(R, R)))
        res[:M] = np.sqrt(np.maximum(nrm2, 0.0)) / np.maximum(np.abs(lambda_[:M]), 1.0)
        epsout.v = float(res[:M].max())
        if epsout.v <= feast_tolerance(fpm) or loop.v >= fpm[3]:
            order = np.argsort(lambda_[:M], kind="stable")                           # feast_sort!
            lambda_[:M] = lambda_[:M][order]
            q[:, :M] = q[:, :M][:, order]
            res[:M] = res[:M][order]
            mode.v = M
            ijob.v = RCI_DONE
            fpm[52] = 0
            state.initialized = False
            return
        loop.v += 1
        Aq[...] = 0
        Sq[...] = 0
        io[:, :M0] = q[:, :M0]
        state.e = 1
        fpm[49] = 1
        state.Q0 = np.array(q[:, :M0], dtype=np.complex128 if hermitian else np.float64)
        Ze.v = complex(state.Zne[0])
        ijob.v = RCI_FACTORIZE
        return
    if ijob.v == RCI_DONE:
        state.initialized = False
        return
    state.initialized = False
    raise ValueError(f"FEAST RCI kernel: Invalid job code ijob={ijob.v}. Expected: -1 (init), 10 (factorize), 11 (solve), "
                     "30 (mult_a), or 0 (done)")


def _custom(contour):
    """(Zne, Wne) of a custom-contour call (with_custom_contour, kernel/feast_kernel.jl:294-336)."""
    Z, W = np.asarray(contour[0], dtype=np.complex128), np.asarray(contour[1], dtype=np.complex128)
    if Z.ndim != 1 or Z.shape != W.shape or len(Z) == 0:
        raise ValueError("custom contour: Zne and Wne must be non-empty vectors of equal length")
    return Z, W


def feast_srci(ijob, N, Ze, work, workc, Aq, Sq, fpm, epsout, loop, Emin, Emax, M0, lambda_, q, mode, res, info, state=None, engine=None,
               contour=None):
    """feast_srci!(ijob, N, Ze, work, workc, Aq, Sq, fpm, epsout, loop, Emin, Emax, M0, lambda, q, mode, res, info; state)
    -- kernel/feast_kernel.jl:7-293 (real symmetric)."""
    state = state if state is not None else FeastRCIState()
    _moment_rci(ijob, N, Ze, work, workc, Aq, Sq, fpm, epsout, loop, Emin, Emax, M0, lambda_, q, mode, res, info, state, False, engine,
                contour)
    return state


def feast_hrci(ijob, N, Ze, work, workc, zAq, zSq, fpm, epsout, loop, Emin, Emax, M0, lambda_, q, mode, res, info, state=None, engine=None,
               contour=None):
    """feast_hrci! -- kernel/feast_kernel.jl:397-644 (complex Hermitian; q, zAq, zSq complex)."""
    state = state if state is not None else FeastRCIState()
    _moment_rci(ijob, N, Ze, work, workc, zAq, zSq, fpm, epsout, loop, Emin, Emax, M0, lambda_, q, mode, res, info, state, True, engine,
                contour)
    return state


def feast_grci(ijob, N, Ze, work, workc, Aq, Sq, fpm, epsout, loop, Emid, r, M0, lambda_, q, mode, res, info, state=None, engine=None,
               contour=None):
    """feast_grci!(ijob, N, Ze, work, workc, Aq, Sq, fpm, epsout, loop, Emid, r, M0, lambda, q, mode, res, info; state)
    -- kernel/feast_kernel.jl:646-962 (general one-sided variant, full contour).  Job sequence per refinement loop:
    10 FACTORIZE -> 11 SOLVE (every node: q += w_e * workc) -> 40 MULT_B (workc <- B q) -> 30 MULT_A (workc <- A q; Aq, Sq formed,
    eigen(Aq, Sq), inside-contour partition, q <- normalised q V) -> 30 MULT_A (workc <- A q[:, :M]; residuals without B) ->
    0 DONE or 10 again.  The block arithmetic runs on the device through the stage-level entry points."""
    from . import feast_gcontour, feast_inside_gcontour, feast_tolerance, feastdefault_
    state = state if state is not None else FeastRCIState()
    if ijob.v == RCI_INIT:
        feastdefault_(fpm)
        info.v = SUCCESS
        if N <= 0:
            info.v = ERR_N
            return state
        if M0 <= 0 or M0 > N:
            info.v = ERR_M0
            return state
        if r <= 0:
            info.v = ERR_EMID_R
            return state
        Z, W = _custom(contour) if contour is not None else feast_gcontour(complex(Emid), float(r), fpm)
        state.Zne, state.Wne, state.ne, state.e, state.initialized = Z.copy(), W.copy(), len(Z), 1, True
        fpm[49], fpm[50], fpm[51], fpm[52] = 1, len(Z), 0, 1
        loop.v = 0
        user = np.array(workc[:, :M0]) if fpm[4] == 1 else None      # fpm[5]: user-provided initial subspace
        for a in (Aq, Sq, lambda_, q, res, workc, work):
            a[...] = 0
        if user is not None:
            for j in range(M0):
                nrm = np.linalg.norm(user[:, j])
                workc[:, j] = user[:, j] / nrm if nrm > 0 else _seed_subspace(N, 1, True)[:, 0]
        else:
            workc[:, :M0] = _seed_subspace(N, M0, True)
        state.Q0 = np.array(workc[:, :M0], dtype=np.complex128)
        state.extra["mult_a_for_projection"] = False
        Ze.v = complex(Z[0])
        ijob.v = RCI_FACTORIZE
        return state
    if ijob.v == RCI_FACTORIZE:
        workc[:, :M0] = state.Q0                                       # kernel/feast_kernel.jl:743-750
        ijob.v = RCI_SOLVE
        return state
    eng = _engine(engine)
    inside = lambda z: bool(feast_inside_gcontour(complex(z), complex(Emid), float(r), fpm))
    if ijob.v == RCI_SOLVE:
        e = state.e
        q[:, :M0] = eng.accumulate(state.Wne[e - 1], np.asarray(workc[:, :M0], dtype=np.complex128), np.asarray(q[:, :M0], dtype=np.complex128))
        state.e = e + 1
        fpm[49] = e + 1
        if e < state.ne:
            Ze.v = complex(state.Zne[e])
            ijob.v = RCI_FACTORIZE
            return state
        state.e = 1
        fpm[49] = 1
        work[...] = 0
        ijob.v = RCI_MULT_B                                            # caller: workc[:, :M0] <- B * q[:, :M0]
        mode.v = M0
        return state
    if ijob.v == RCI_MULT_B:
        Sq[:M0, :M0] = eng.gram(np.asarray(q[:, :M0], dtype=np.complex128), np.asarray(workc[:, :M0], dtype=np.complex128))   # q^H (B q)
        workc[...] = 0
        ijob.v = RCI_MULT_A                                            # caller: workc[:, :M0] <- A * q[:, :M0]
        mode.v = M0
        state.extra["mult_a_for_projection"] = True
        return state
    if ijob.v == RCI_MULT_A:
        if state.extra.get("mult_a_for_projection"):
            state.extra["mult_a_for_projection"] = False
            Aq[:M0, :M0] = eng.gram(np.asarray(q[:, :M0], dtype=np.complex128), np.asarray(workc[:, :M0], dtype=np.complex128))  # q^H (A q)
            try:
                lam_red, V = eng.eig_general(np.asarray(Aq[:M0, :M0], dtype=np.complex128), np.asarray(Sq[:M0, :M0], dtype=np.complex128))
            except Exception as err:                                   # kernel/feast_kernel.jl:885-891 catch
                state.extra["error"] = str(err)
                info.v = ERR_LAPACK
                ijob.v = RCI_DONE
                fpm[52] = 0
                state.initialized = False
                return state
            flags = [bool(np.isfinite(z)) and inside(z) for z in lam_red]
            perm = [i for i in range(M0) if flags[i]] + [i for i in range(M0) if not flags[i]]
            M = sum(flags)
            fpm[51] = M
            state.M = M
            if M == 0:
                info.v = ERR_NO_CONV
                ijob.v = RCI_DONE
                fpm[52] = 0
                state.initialized = False
                return state
            X = eng.rowtransform(np.asarray(q[:, :M0], dtype=np.complex128), np.asarray(V, dtype=np.complex128))[:, perm]   # q V, all M0
            nrm2 = np.real(np.diag(eng.gram(X, X)))
            scale = np.where(nrm2 > 0, 1.0 / np.sqrt(np.where(nrm2 > 0, nrm2, 1.0)), 1.0)
            q[:, :M0] = eng.rowtransform(X, np.diag(scale).astype(np.complex128))                   # unit 2-norm columns
            lambda_[:M0] = np.asarray(lam_red)[perm]
            workc[...] = 0
            ijob.v = RCI_MULT_A                                        # caller: workc[:, :M] <- A * q[:, :M] (residuals)
            mode.v = M
            return state
        M = fpm[51]
        Lq = eng.rowtransform(np.asarray(q[:, :M], dtype=np.complex128), np.diag(lambda_[:M]).astype(np.complex128))
        R = eng.accumulate(-1.0, Lq, np.asarray(workc[:, :M], dtype=np.complex128))
        nrm2 = np.real(np.diag(eng.gram(R, R)))
        res[:M] = np.sqrt(np.maximum(nrm2, 0.0)) / np.maximum(np.abs(lambda_[:M]), 1.0)      # B is not part of it (:900-906)
        epsout.v = float(res[:M].max())
        if epsout.v <= feast_tolerance(fpm) or loop.v >= fpm[3]:
            order = np.argsort(np.abs(lambda_[:M]), kind="stable")                            # feast_sort_general!: by |lambda|
            lambda_[:M] = lambda_[:M][order]
            q[:, :M] = q[:, :M][:, order]
            res[:M] = res[:M][order]
            mode.v = M
            ijob.v = RCI_DONE
            fpm[52] = 0
            state.initialized = False
            return state
        loop.v += 1
        state.Q0 = np.array(q[:, :M0], dtype=np.complex128)
        Aq[...] = 0
        Sq[...] = 0
        q[...] = 0
        workc[:, :M0] = state.Q0
        Z, W = _custom(contour) if contour is not None else feast_gcontour(complex(Emid), float(r), fpm)
        state.Zne, state.Wne, state.ne, state.e = Z.copy(), W.copy(), len(Z), 1
        fpm[49] = 1
        Ze.v = complex(Z[0])
        ijob.v = RCI_FACTORIZE
        return state
    if ijob.v == RCI_DONE:
        state.initialized = False
        return state
    state.initialized = False
    raise ValueError(f"FEAST RCI kernel (General): Invalid job code ijob={ijob.v}. Expected: -1 (init), 10 (factorize), 11 (solve), "
                     "40 (mult_b), 30 (mult_a), or 0 (done)")


# custom-contour forms (kernel/feast_kernel.jl:294-336): the caller's nodes and weights replace feast_contour / feast_gcontour
def feast_srcix(ijob, N, Ze, work, workc, Aq, Sq, fpm, epsout, loop, Emin, Emax, M0, lambda_, q, mode, res, info, Zne, Wne, **kw):
    return feast_srci(ijob, N, Ze, work, workc, Aq, Sq, fpm, epsout, loop, Emin, Emax, M0, lambda_, q, mode, res, info, contour=(Zne, Wne), **kw)


def feast_hrcix(ijob, N, Ze, work, workc, zAq, zSq, fpm, epsout, loop, Emin, Emax, M0, lambda_, q, mode, res, info, Zne, Wne, **kw):
    return feast_hrci(ijob, N, Ze, work, workc, zAq, zSq, fpm, epsout, loop, Emin, Emax, M0, lambda_, q, mode, res, info, contour=(Zne, Wne), **kw)


def feast_grcix(ijob, N, Ze, work, workc, Aq, Sq, fpm, epsout, loop, Emid, r, M0, lambda_, q, mode, res, info, Zne, Wne, **kw):
    return feast_grci(ijob, N, Ze, work, workc, Aq, Sq, fpm, epsout, loop, Emid, r, M0, lambda_, q, mode, res, info, contour=(Zne, Wne), **kw)


# parallel alias of the real RCI kernel (interfaces/feast_precision_aliases.jl + parallel/feast_parallel_rci.jl:47-266) and the
# solver-neutral "iterative FEAST" names (kernel/feast_kernel.jl:346-395): same state machines
def pdfeast_srci(*a, **kw):
    return feast_srci(*a, **kw)


dfeast_srci = feast_srci
zfeast_hrci = feast_hrci
zfeast_grci = feast_grci
ifeast_srci = feast_srci
ifeast_hrci = feast_hrci
ifeast_grci = feast_grci
