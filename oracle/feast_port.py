"""CPU port of the ENGINE's algorithm (test infrastructure, like everything under oracle/).

``feast_oracle.py`` restates the reference (direct solves / per-column GMRES).  The CUDA
engine replaces the per-node solve by a lock-step multi-RHS BiCGStab with an optional
Ritz-pair initial guess and inexact ("IFEAST"-style) stopping; this file is the NumPy/SciPy
port of exactly that variant.  It is used
  * by tests, to compare the CUDA block solver / refinement loop against the same
    arithmetic on the CPU, and
  * by bench.py as the ``cpu_baseline`` (kind "port") and ``--impl reference`` arm, because the
    reference's own iterative path (restarted GMRES(30), 500 iterations, tol 1e-12,
    sparse/feast_sparse.jl:164-236) cannot converge on the n=1e6 Laplacian and its direct path
    (UMFPACK, sparse/feast_sparse.jl:339) does not fit in memory there.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import it.
"""
from __future__ import annotations

import math
import time

import numpy as np
import scipy.sparse as sp

import feast_oracle as fo


def shifted_apply(A, B, z, X):
    """Y = z*(B X) - A X  (sparse/feast_sparse.jl:20-27, 142-148)."""
    return z * (X if B is None else B @ X) - A @ X


def block_bicgstab(A, B, z, RHS, X0=None, rtol=1e-12, atol=None, maxiter=500, rel_to_initial=0.0,
                   max_restarts=3, stats=None):
    """Lock-step multi-RHS BiCGStab for (zB-A)X = RHS, per-column scalars and masks.

    Stopping rule per column (Krylov.jl convention used at sparse/feast_sparse.jl:183-188):
    ||r|| <= atol + rtol*||b||; additionally, when rel_to_initial > 0, a column also stops once
    ||r|| <= rel_to_initial*||r_0|| (inexact / IFEAST mode, r_0 = residual of the initial guess).
    After the recursive residual reports convergence the TRUE residual is recomputed and the
    iteration restarted from it when it disagrees (at most max_restarts times).
    Returns (X, iters_per_column, true_residual_norms, converged_mask).
    """
    n, m = RHS.shape
    atol = rtol if atol is None else atol
    X = np.zeros((n, m), dtype=np.complex128) if X0 is None else np.array(X0, dtype=np.complex128)
    bn = np.linalg.norm(RHS, axis=0)
    target = atol + rtol * bn
    its = np.zeros(m, dtype=np.int64)
    first = True
    tiny = np.finfo(float).tiny
    for _restart in range(max_restarts + 1):
        R = RHS - shifted_apply(A, B, z, X)
        rn = np.linalg.norm(R, axis=0)
        if first:
            if rel_to_initial > 0:
                target = np.maximum(target, rel_to_initial * rn)
            first = False
        active = rn > target
        if not active.any() or its.max() >= maxiter:
            break
        Rh = R.copy()
        rho = np.sum(Rh.conj() * R, axis=0)
        P = R.copy()
        Xb, rn0 = X.copy(), rn.copy()
        while active.any() and its.max() < maxiter:
            V = shifted_apply(A, B, z, P)
            den = np.sum(Rh.conj() * V, axis=0)
            ok = active & (np.abs(den) > tiny)
            alpha = np.where(ok, rho / np.where(ok, den, 1), 0)
            S = R - alpha * V
            T = shifted_apply(A, B, z, S)
            tt = np.sum(np.abs(T) ** 2, axis=0)
            ts = np.sum(T.conj() * S, axis=0)
            ok2 = ok & (tt > tiny)
            omega = np.where(ok2, ts / np.where(ok2, tt, 1), 0)
            ss = np.sum(np.abs(S) ** 2, axis=0)
            with np.errstate(divide="ignore", invalid="ignore"):
                cosang = np.where(ok2 & (ss > tiny), np.abs(ts) / np.sqrt(np.where(ok2, tt, 1) * np.where(ss > tiny, ss, 1)), 1.0)
            omega = np.where((cosang > 0) & (cosang < 0.7), omega * (0.7 / np.where(cosang > 0, cosang, 1)), omega)
            rht = np.sum(Rh.conj() * T, axis=0)
            rho_new = -omega * rht
            ok3 = ok2 & (np.abs(omega) > tiny) & (np.abs(rho) > tiny)
            beta = np.where(ok3, (rho_new / np.where(ok3, rho, 1)) * (alpha / np.where(ok3, omega, 1)), 0)
            X += alpha * P + omega * S
            R = S - omega * T
            P = R + beta * (P - omega * V)
            rho = np.where(ok3, rho_new, rho)
            its[active] += 1
            rn = np.linalg.norm(R, axis=0)
            div = active & ~(rn <= 1e4 * rn0)
            if div.any():
                X[:, div] = Xb[:, div]
            active = active & ok3 & (rn > target) & ~div
            if stats is not None:
                stats["spmm"] = stats.get("spmm", 0) + 2
    true = np.linalg.norm(RHS - shifted_apply(A, B, z, X), axis=0)
    if stats is not None:
        stats["krylov_iters"] = stats.get("krylov_iters", 0) + int(its.max())
        stats["col_iters"] = stats.get("col_iters", 0) + int(its.sum())
    return X, its, true, true <= 10 * np.maximum(target, 0)


def feast_hrr_bicgstab(A, B, Emin, Emax, M0, fpm, Q0, inner_rtol=None, inner_rel=0.0, inner_maxiter=500,
                       ritz_guess=True, filter="true", verbose=False, node_items=None, allreduce=None):
    """H-RR refinement loop (dense/feast_dense.jl:78-351 skeleton) with the engine's block solver.

    node_items: optional list of (node, col0, ncols) work items owned by this rank (multi-GPU
    emulation); allreduce: callable summing the accumulator over ranks.
    """
    N = A.shape[0]
    fo.feastdefault(fpm)
    fo.check_feast_srci_input(N, M0, Emin, Emax, fpm)
    tol_value = 10.0 ** (-fpm[2]) if inner_rtol is None else inner_rtol
    real_mode = filter == "true" and not np.iscomplexobj(A) and (B is None or not np.iscomplexobj(B)) \
        and not np.iscomplexobj(Q0)
    Zne, Wne = fo.feast_contour(Emin, Emax, fpm)
    ne = len(Zne)
    Qb = np.array(Q0, dtype=np.complex128)
    maxloop = fpm[3]
    eps_tol = fo.feast_tolerance(fpm)
    lam = np.zeros(M0)
    res = np.zeros(M0)
    X = np.zeros((N, M0), dtype=np.complex128)
    have_ritz = False
    active = M0
    info, epsout, loop_count, M_found = fo.SUCCESS, math.inf, 0, 0
    stats = {"krylov_iters": 0, "col_iters": 0, "spmm": 0, "node_iters": []}
    for loop_idx in range(maxloop + 1):
        loop_count = loop_idx
        acc = np.zeros((N, M0), dtype=np.complex128)
        basis = Qb[:, :active]
        rhs = basis if B is None else B @ basis
        items = node_items(loop_idx, active) if node_items is not None else [(e, 0, active) for e in range(ne)]
        per_node = []
        for (e, c0, nc) in items:
            z = Zne[e]
            X0 = None
            if ritz_guess and have_ritz:
                X0 = basis[:, c0:c0 + nc] / (z - lam[c0:c0 + nc])
            Y, its, true, ok = block_bicgstab(A, B, z, rhs[:, c0:c0 + nc], X0, rtol=tol_value, maxiter=inner_maxiter,
                                              rel_to_initial=inner_rel, stats=stats)
            per_node.append(int(its.max()))
            acc[:, c0:c0 + nc] += 2 * Wne[e] * Y
        stats["node_iters"].append(per_node)
        if allreduce is not None:
            acc = allreduce(acc)
        if real_mode:
            acc = acc.real.astype(np.complex128)
            Qr, rank = fo.qr_compress(np.ascontiguousarray(acc.real), active)
            Qr = Qr.astype(np.complex128)
        else:
            Qr, rank = fo.qr_compress(acc, active)
        if rank == 0:
            info = fo.ERR_NO_CONV
            break
        Sq = fo.hermitian_part(Qr.conj().T @ (A @ Qr))
        Aq = np.eye(rank, dtype=np.complex128) if B is None else fo.hermitian_part(Qr.conj().T @ (B @ Qr))
        import scipy.linalg as sla
        lam_red, v_red = sla.eigh(Sq.real, Aq.real) if real_mode else sla.eigh(Sq, Aq)
        X[:, :rank] = Qr @ v_red
        lam[:rank] = lam_red
        M = fo.reorder_by_interval(lam, X, Emin, Emax, rank)
        if M == 0:
            info = fo.ERR_NO_CONV
            break
        X[:, :M] /= np.linalg.norm(X[:, :M], axis=0)
        for j in range(M):
            x = X[:, j]
            rv = A @ x - lam[j] * (x if B is None else B @ x)
            res[j] = np.linalg.norm(rv) / max(abs(lam[j]), 1.0)
        epsout = float(res[:M].max())
        M_found = M
        if verbose:
            print(f"loop {loop_idx}: M={M} rank={rank} epsout={epsout:.3e} node_iters={per_node}", flush=True)
        if epsout <= eps_tol:
            break
        if loop_idx == maxloop:
            info = fo.ERR_NO_CONV
            break
        active = rank
        Qb[:, :active] = X[:, :active]
        have_ritz = True
    return fo.FeastResult(lam[:M_found].copy(), X[:, :M_found].copy(), M_found, res[:M_found].copy(), info, epsout,
                          loop_count, stats)


def time_bicgstab_sample(A, z, m, iters, seed=0):
    """CPU-baseline sample: `iters` lock-step BiCGStab iterations on m columns; returns seconds."""
    n = A.shape[0]
    rng = np.random.default_rng(seed)
    RHS = rng.standard_normal((n, m)).astype(np.complex128)
    t0 = time.perf_counter()
    block_bicgstab(A, None, z, RHS, None, rtol=0.0, atol=0.0, maxiter=iters, max_restarts=0)
    return time.perf_counter() - t0
