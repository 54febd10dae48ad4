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


# --------------------------------------------------------------------------------------------------
# multi-shift two-pass Lanczos (the engine's solver for standard real-symmetric problems)
# --------------------------------------------------------------------------------------------------
_A32_CACHE = {}


def _narrow_operator(A):
    """FP32 copy of the operator's entries (k_lz32_vals), cached for the LAST matrix object (the entry keeps a reference to
    it: an id() alone can be reused by a different matrix after garbage collection)."""
    ent = _A32_CACHE.get("last")
    if ent is None or ent[0] is not A:
        ent = (A, A.astype(np.float32))
        _A32_CACHE["last"] = ent
    return ent[1]


def lanczos_pass1(A, b, Zne, kmax, target, mixed=False):
    """Lock-step Lanczos on the columns of the real block b; stops after the first step k at which every
    shifted system's residual estimate beta_{k+1} |e_k^T (z_e I - T_k)^-1 e_1| (relative to ||b||) is <= target.
    Mirrors kernels_lanczos.cuh: k_lz_spmm<LZ_P1> / k_lz_update / k_lz_scal1 / k_lz_scal2 (unnormalised vectors,
    breakdown freeze at beta <= 1e-13 * max(|alpha|, beta)).  Returns alpha (k x m), beta ((k+1) x m), k, maxres.
    mixed: kernels_lanczos_f32.cuh -- FP32 vectors/entries/applied scalars, FP64 dot products and recurrences."""
    n, m = b.shape
    vdt = np.float32 if mixed else b.dtype
    Aop = _narrow_operator(A) if mixed else A
    rnd = (lambda x: np.asarray(x, dtype=np.float32)) if mixed else (lambda x: x)
    ne = len(Zne)
    alpha = np.zeros((kmax, m))
    beta = np.zeros((kmax + 1, m))
    beta[0] = np.linalg.norm(b, axis=0)
    inv = np.where(beta[0] > 1e-290, 1.0 / np.where(beta[0] > 0, beta[0], 1.0), 0.0)
    scale = np.zeros(m)
    u = b.astype(vdt)
    u_prev = np.zeros_like(u)
    ratio_b = np.zeros(m)
    d = np.zeros((ne, m), dtype=complex)
    g = np.zeros((ne, m), dtype=complex)
    k, maxres = 0, math.inf
    acc_dt = np.float64 if mixed else None
    for j in range(kmax):
        t = (Aop @ u) * rnd(inv) - rnd(ratio_b) * u_prev
        al = np.einsum("ij,ij->j", u.conj(), t, dtype=acc_dt).real * inv     # real for Hermitian A (complex vectors allowed)
        alpha[j] = al
        scale = np.maximum(scale, np.abs(al))
        u_next = t - rnd(al * inv) * u
        bn = np.sqrt(np.einsum("ij,ij->j", u_next.conj(), u_next, dtype=acc_dt).real)
        ok = (bn > 1e-290) & (bn > 1e-13 * scale) & (inv != 0.0)
        beta[j + 1] = np.where(ok, bn, 0.0)
        inv_next = np.where(ok, 1.0 / np.where(bn > 0, bn, 1.0), 0.0)
        ratio_b = np.where(ok, bn * inv, 0.0)
        scale = np.maximum(scale, bn)
        for e in range(ne):
            if j == 0:
                d[e] = Zne[e] - al
                g[e] = 1.0 / d[e]
            else:
                dn = (Zne[e] - al) - beta[j] ** 2 / d[e]
                g[e] = beta[j] * g[e] / dn
                d[e] = dn
        maxres = float((beta[j + 1][None, :] * np.abs(g)).max())
        u_prev, u, inv = u, u_next, inv_next
        k = j + 1
        if maxres <= target:
            break
    return alpha[:k], beta[:k + 1], k, maxres


def lanczos_coefficients(alpha, beta, Zne, Wne, F):
    """c_j = ||b|| sum_e Re(2 w_e F_e [(z_e I - T_k)^-1 e_1]_j) / beta_j  (k x m); F: ne x m guess factors."""
    k, m = alpha.shape
    coef = np.zeros((k, m))
    for c in range(m):
        if not beta[0, c] > 0:
            continue
        kc = k
        for j in range(1, k):
            if beta[j, c] == 0.0:
                kc = j
                break
        for e, z in enumerate(Zne):
            ab = np.zeros((3, kc), dtype=complex)
            ab[1] = z - alpha[:kc, c]
            ab[0, 1:] = -beta[1:kc, c]
            ab[2, :-1] = -beta[1:kc, c]
            rhs = np.zeros(kc, dtype=complex)
            rhs[0] = 1.0
            import scipy.linalg as sla
            y = sla.solve_banded((1, 1), ab, rhs)
            coef[:kc, c] += np.real(2 * Wne[e] * F[e, c] * beta[0, c] * y)
        coef[:kc, c] /= beta[:kc, c]
    return coef


def lanczos_pass2(A, b, alpha, beta, coef, Q, mixed=False):
    """Re-run the recurrence with the stored scalars and accumulate Q += coef_j * u_j (k_lz_spmm<LZ_P2>; mixed:
    k_lz32_spmm<LZ_P2>, FP32 vectors into the FP64 accumulator)."""
    k, m = alpha.shape
    inv = np.where(beta > 0, 1.0 / np.where(beta > 0, beta, 1.0), 0.0)
    vdt = np.float32 if mixed else b.dtype
    Aop = _narrow_operator(A) if mixed else A
    rnd = (lambda x: np.asarray(x, dtype=np.float32)) if mixed else (lambda x: x)
    u = b.astype(vdt)
    u_prev = np.zeros_like(u)
    for j in range(k):
        Q += coef[j] * u
        if j == k - 1:
            break
        ratio_b = beta[j] * inv[j - 1] if j > 0 else np.zeros(m)
        t = (Aop @ u) * rnd(inv[j]) - rnd(ratio_b) * u_prev
        u_prev, u = u, t - rnd(alpha[j] * inv[j]) * u
    return Q


def mslanczos_filter(A, Q, theta, Zne, Wne, target, kmax, stats=None, mixed=False):
    """sum_e Re(2 w_e (z_e I - A)^-1 q) for the real block Q; theta (Ritz values) or None (zero guess)."""
    n, m = Q.shape
    ne = len(Zne)
    if theta is None:
        b = Q
        F = np.ones((ne, m), dtype=complex)
        acc = np.zeros((n, m), dtype=Q.dtype)
    else:
        b = A @ Q - Q * theta
        F = 1.0 / (np.asarray(Zne)[:, None] - theta[None, :])
        acc = Q * np.real((2 * np.asarray(Wne)[:, None] * F).sum(axis=0))
    mixed = mixed and not np.iscomplexobj(b)
    alpha, beta, k, maxres = lanczos_pass1(A, b, Zne, kmax, target, mixed)
    coef = lanczos_coefficients(alpha, beta, Zne, Wne, F)
    acc = lanczos_pass2(A, b, alpha, beta, coef, acc, mixed)
    if stats is not None:
        stats["lz_steps"].append(k)
        stats["lz_maxres"].append(maxres)
    return acc


def feast_hrr_mslanczos(A, Emin, Emax, M0, fpm, Q0, inner_rel=1e-3, inner_rel0=0.0, inner_maxiter=500, ritz_guess=True,
                        verbose=False, col_slices=None, allreduce=None, adaptive=False, mixed=False):
    """H-RR refinement loop (sparse/feast_sparse.jl:246-499 skeleton, B = I, real symmetric A, real Q0, filter rho = Re g)
    with the engine's multi-shift Lanczos inner solver.  col_slices(active) -> (c0, nc) emulates one rank of the
    column-sharded multi-GPU run; allreduce sums the accumulator over ranks.  mixed: FP32 Lanczos vectors (fpm[42]) until a
    refined sweep gains less than a factor 4, then FP64 -- run_interval's rule."""
    import scipy.linalg as sla
    N = A.shape[0]
    fo.feastdefault(fpm)
    fo.check_feast_srci_input(N, M0, Emin, Emax, fpm)
    tol_value = 10.0 ** (-fpm[2])
    Zne, Wne = fo.feast_contour(Emin, Emax, fpm)
    cplx = np.iscomplexobj(A) or np.iscomplexobj(Q0)     # complex Hermitian A: complex vectors, real tridiagonal and coefficients
    wdt = np.complex128 if cplx else np.float64
    Qb = np.array(Q0, dtype=wdt)
    maxloop = fpm[3]
    eps_tol = fo.feast_tolerance(fpm)
    lam = np.zeros(M0)
    res = np.zeros(M0)
    X = np.zeros((N, M0), dtype=wdt)
    have_ritz = False
    active = M0
    info, epsout, loop_count, M_found = fo.SUCCESS, math.inf, 0, 0
    stats = {"lz_steps": [], "lz_maxres": [], "fp32_sweeps": 0}
    use_fp32, eps_before = bool(mixed) and not cplx, math.inf
    for loop_idx in range(maxloop + 1):
        loop_count = loop_idx
        first = not (ritz_guess and have_ritz)
        target = inner_rel0 if (first and inner_rel0 > 0) else inner_rel
        if not target > 0:
            target = tol_value
        if adaptive and not first and math.isfinite(epsout) and epsout > 0:
            t = 2.0 * eps_tol / epsout      # run_interval in csrc/feastcuda.cu: aim the sweep at the tolerance when in reach
            if t >= 1e-6:
                target = min(0.1, t)
        c0, nc = (0, active) if col_slices is None else col_slices(active)
        acc = np.zeros((N, active), dtype=wdt)
        if nc > 0:
            acc[:, c0:c0 + nc] = mslanczos_filter(A, Qb[:, c0:c0 + nc], None if first else lam[c0:c0 + nc], Zne, Wne, target,
                                                  inner_maxiter, stats, use_fp32)
            stats["fp32_sweeps"] += int(use_fp32)
        if allreduce is not None:
            acc = allreduce(acc)
        Qr, rank = fo.qr_compress(np.ascontiguousarray(acc), active)
        if rank == 0:
            info = fo.ERR_NO_CONV
            break
        if not cplx:
            Qr = Qr.real
        Sq = Qr.conj().T @ (A @ Qr)
        Sq = 0.5 * (Sq + Sq.conj().T)
        lam_red, v_red = sla.eigh(Sq)
        Xc = np.zeros((N, M0), dtype=np.complex128)
        Xc[:, :rank] = Qr @ v_red
        lam[:rank] = lam_red
        M = fo.reorder_by_interval(lam, Xc, Emin, Emax, rank)
        X = Xc.copy() if cplx else Xc.real.copy()
        if M == 0:
            info = fo.ERR_NO_CONV
            break
        X[:, :M] /= np.linalg.norm(X[:, :M], axis=0)
        R = A @ X[:, :M] - X[:, :M] * lam[:M]
        res[:M] = np.linalg.norm(R, axis=0) / np.maximum(np.abs(lam[:M]), 1.0)
        epsout = float(res[:M].max())
        M_found = M
        if use_fp32 and have_ritz and math.isfinite(eps_before) and not epsout <= 0.25 * eps_before:
            use_fp32 = False
        eps_before = epsout
        if verbose:
            print(f"loop {loop_idx}: M={M} rank={rank} epsout={epsout:.3e} k={stats['lz_steps'][-1] if stats['lz_steps'] else 0}", flush=True)
        if epsout <= eps_tol:
            break
        if loop_idx == maxloop:
            info = fo.ERR_NO_CONV
            break
        active = rank
        Qb = X[:, :active].copy()
        have_ritz = True
    return fo.FeastResult(lam[:M_found].copy(), X[:, :M_found].copy(), M_found, res[:M_found].copy(), info, epsout,
                          loop_count, stats)


_POOL_A = None


def _sample_worker(args):
    """One worker = one slice of columns, like one thread of the reference's threaded backend (parallel/feast_parallel.jl:586)."""
    m, steps, seed = args
    A = _POOL_A
    n = A.shape[0]
    rng = np.random.default_rng(seed)
    b = rng.standard_normal((n, m))
    Z = np.array([0.5 + 0.5j])
    t0 = time.perf_counter()
    alpha, beta, k, _ = lanczos_pass1(A, b, Z, steps, 0.0)
    t1 = time.perf_counter()
    coef = np.ones((k, m))
    lanczos_pass2(A, b, alpha, beta, coef, np.zeros((n, m)))
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1


def time_mslanczos_sample(A, m, steps, workers=1, seed=0):
    """CPU-baseline sample: `steps` lock-step Lanczos steps of pass 1 and of pass 2 on m real columns, the columns split
    over `workers` processes.  Returns (seconds pass 1, seconds pass 2) = wall time of the slowest worker."""
    global _POOL_A
    _POOL_A = A
    workers = max(1, min(workers, m))
    base, rem = divmod(m, workers)
    jobs = [(base + (1 if w < rem else 0), steps, seed + w) for w in range(workers)]
    if workers == 1:
        out = [_sample_worker(jobs[0])]
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(workers) as pool:
            t0 = time.perf_counter()
            out = pool.map(_sample_worker, jobs)
            wall = time.perf_counter() - t0
        tot = max(a + b for a, b in out)
        scale = wall / tot if tot > 0 else 1.0      # include pool overheads in the reported wall time
        out = [(a * scale, b * scale) for a, b in out]
    return max(o[0] for o in out), max(o[1] for o in out)


def time_bicgstab_sample(A, z, m, iters, seed=0):
    """CPU-baseline sample: `iters` lock-step BiCGStab iterations on m columns; returns seconds."""
    n = A.shape[0]
    rng = np.random.default_rng(seed)
    RHS = rng.standard_normal((n, m)).astype(np.complex128)
    t0 = time.perf_counter()
    block_bicgstab(A, None, z, RHS, None, rtol=0.0, atol=0.0, maxiter=iters, max_restarts=0)
    return time.perf_counter() - t0


# ======================================================================================================================
# Row-sharded variant of the multi-shift Lanczos solve -- the multi-GPU scheme planned in DESIGN.md §8 (item 3), restated
# on the CPU so that its arithmetic, its communication pattern and its parity with the single-rank solve can be tested
# before any CUDA exists for it.  Every rank owns the rows [r0, r1) of A and of every block vector:
#   * SpMM: the rank needs the rows of u its stored columns reference ("halo"); `exchange(local_block)` returns the block
#     of ALL rows (an all-gather here; NVLink peer loads of the halo rows only in the CUDA design);
#   * the two dot products of a pass-1 step become `allreduce_small` calls on m numbers; pass 2 needs none;
#   * the Rayleigh-Ritz stage acts on local rows: Gram matrices are summed with `allreduce_small`, row transforms are local.
# Orthonormalisation: CholQR2 through the eigen-decomposition of the Gram matrix with the rank threshold of
# _feast_qr_compress! (core/feast_aux.jl:101-131) applied to the singular values sqrt(eig(G)).
# ======================================================================================================================
def _rows_orthonormalize(acc_loc, allreduce_small, n_total, rank_tol=None):
    m = acc_loc.shape[1]
    eps = np.finfo(np.float64).eps
    rank_tol = math.sqrt(eps) if rank_tol is None else rank_tol
    G = allreduce_small(acc_loc.conj().T @ acc_loc)
    w, V = np.linalg.eigh(0.5 * (G + G.conj().T))
    sv = np.sqrt(np.maximum(w, 0.0))
    if not sv.max() > 0:
        return acc_loc[:, :0].copy(), 0
    keep = sv > max(rank_tol, eps * max(n_total, m)) * sv.max()
    rank = int(keep.sum())
    Q = acc_loc @ (V[:, keep] / sv[keep])
    for _ in range(2):                                   # re-orthonormalise: CholQR2
        G2 = allreduce_small(Q.conj().T @ Q)
        Lc = np.linalg.cholesky(0.5 * (G2 + G2.conj().T))
        Q = np.linalg.solve(Lc, Q.conj().T).conj().T
    return Q, rank


def feast_hrr_mslanczos_rows(A, Emin, Emax, M0, fpm, Q0, r0, r1, exchange, allreduce_small, inner_rel=1e-3, inner_maxiter=500,
                             adaptive=True, verbose=False):
    """One rank of the row-sharded solve (real symmetric A, B = I).  Returns the FeastResult with THIS rank's rows of q."""
    N = A.shape[0]
    fo.feastdefault(fpm)
    fo.check_feast_srci_input(N, M0, Emin, Emax, fpm)
    Zne, Wne = fo.feast_contour(Emin, Emax, fpm)
    ne = len(Zne)
    eps_tol = fo.feast_tolerance(fpm)
    A_loc = A[r0:r1].tocsr()
    dots = lambda X, Y: allreduce_small(np.einsum("ij,ij->j", X, Y))
    apply_rows = lambda X_loc: A_loc @ exchange(X_loc)

    def filter_rows(Q_loc, theta, target):
        m = Q_loc.shape[1]
        if theta is None:
            b, F, acc = Q_loc.copy(), np.ones((ne, m), dtype=complex), np.zeros_like(Q_loc)
        else:
            b = apply_rows(Q_loc) - Q_loc * theta
            F = 1.0 / (np.asarray(Zne)[:, None] - theta[None, :])
            acc = Q_loc * np.real((2 * np.asarray(Wne)[:, None] * F).sum(axis=0))
        # pass 1 (lanczos_pass1 with distributed dots)
        alpha, beta = np.zeros((inner_maxiter, m)), np.zeros((inner_maxiter + 1, m))
        beta[0] = np.sqrt(dots(b, b))
        inv = np.where(beta[0] > 1e-290, 1.0 / np.where(beta[0] > 0, beta[0], 1.0), 0.0)
        scale, ratio_b = np.zeros(m), np.zeros(m)
        u_prev, u = np.zeros_like(b), b.copy()
        d, g = np.zeros((ne, m), dtype=complex), np.zeros((ne, m), dtype=complex)
        k = 0
        for j in range(inner_maxiter):
            t = apply_rows(u) * inv - ratio_b * u_prev
            al = dots(u, t) * inv
            alpha[j] = al
            scale = np.maximum(scale, np.abs(al))
            u_next = t - (al * inv) * u
            bn = np.sqrt(dots(u_next, u_next))
            ok = (bn > 1e-290) & (bn > 1e-13 * scale) & (inv != 0.0)
            beta[j + 1] = np.where(ok, bn, 0.0)
            inv_next = np.where(ok, 1.0 / np.where(bn > 0, bn, 1.0), 0.0)
            ratio_b = np.where(ok, bn * inv, 0.0)
            scale = np.maximum(scale, bn)
            for e in range(ne):
                if j == 0:
                    d[e] = Zne[e] - al
                    g[e] = 1.0 / d[e]
                else:
                    dn = (Zne[e] - al) - beta[j] ** 2 / d[e]
                    g[e] = beta[j] * g[e] / dn
                    d[e] = dn
            maxres = float((beta[j + 1][None, :] * np.abs(g)).max())
            u_prev, u, inv = u, u_next, inv_next
            k = j + 1
            if maxres <= target:
                break
        alpha, beta = alpha[:k], beta[:k + 1]
        coef = lanczos_coefficients(alpha, beta, Zne, Wne, F)        # replicated: every rank holds the same scalars
        # pass 2: no reductions at all
        invb = np.where(beta > 0, 1.0 / np.where(beta > 0, beta, 1.0), 0.0)
        u_prev, u = np.zeros_like(b), b.copy()
        for j in range(k):
            acc += coef[j] * u
            if j == k - 1:
                break
            rb = beta[j] * invb[j - 1] if j > 0 else np.zeros(m)
            t = apply_rows(u) * invb[j] - rb * u_prev
            u_prev, u = u, t - (alpha[j] * invb[j]) * u
        return acc, k

    Qb = np.array(Q0[r0:r1], dtype=np.float64)
    lam, res = np.zeros(M0), np.zeros(M0)
    X = np.zeros((r1 - r0, M0))
    have_ritz, active = False, M0
    info, epsout, loop_count, M_found = fo.SUCCESS, math.inf, 0, 0
    stats = {"lz_steps": [], "exchanges": 0, "small_allreduces": 0}
    for loop_idx in range(fpm[3] + 1):
        loop_count = loop_idx
        target = inner_rel
        if adaptive and have_ritz and math.isfinite(epsout) and epsout > 0:
            t = 2.0 * eps_tol / epsout
            if t >= 1e-6:
                target = min(0.1, t)
        acc, k = filter_rows(Qb[:, :active], lam[:active].copy() if have_ritz else None, target)
        stats["lz_steps"].append(k)
        Qr, rank = _rows_orthonormalize(acc, allreduce_small, N)
        if rank == 0:
            info = fo.ERR_NO_CONV
            break
        Sq = allreduce_small(Qr.T @ apply_rows(Qr))
        lam_red, v_red = np.linalg.eigh(0.5 * (Sq + Sq.T))
        Xc = np.zeros((r1 - r0, M0), dtype=complex)
        Xc[:, :rank] = Qr @ v_red
        lam[:rank] = lam_red
        M = fo.reorder_by_interval(lam, Xc, Emin, Emax, rank)
        X = Xc.real.copy()
        if M == 0:
            info = fo.ERR_NO_CONV
            break
        X[:, :M] /= np.sqrt(dots(X[:, :M], X[:, :M]))
        R = apply_rows(X[:, :M]) - X[:, :M] * lam[:M]
        res[:M] = np.sqrt(dots(R, R)) / np.maximum(np.abs(lam[:M]), 1.0)
        epsout = float(res[:M].max())
        M_found = M
        if verbose:
            print(f"rows[{r0}:{r1}] loop {loop_idx}: M={M} rank={rank} epsout={epsout:.3e} k={k}", flush=True)
        if epsout <= eps_tol:
            break
        if loop_idx == fpm[3]:
            info = fo.ERR_NO_CONV
            break
        active = rank
        Qb = X[:, :active].copy()
        have_ritz = True
    return fo.FeastResult(lam[:M_found].copy(), X[:, :M_found].copy(), M_found, res[:M_found].copy(), info, epsout, loop_count, stats)


# ======================================================================================================================
# Generalized Hermitian problems A x = lambda B x (B Hermitian positive definite) with the multi-shift Lanczos filter --
# the design for BASELINE configs[3], restated on the CPU before any CUDA exists for it.
# C = B^-1 A is self-adjoint in the B-inner product <x, y>_B = x^H B y, and (z B - A)^-1 B q = (z I - C)^-1 q, so ONE Lanczos
# recurrence in that inner product serves every node exactly as in the standard case: T_k is real, the true filter
# rho(C) q = V_k rho(T_k) e_1 ||q||_B has real coefficients, no adjoint solves are needed.  A step costs one product with A
# and one solve with B (mass matrices are well conditioned: Jacobi-preconditioned CG, a fixed and small number of products
# with B); the B-images p_j = B v_j ride along so that no extra product with B is needed for the norms:
#     r = A v_j - alpha_j p_j - beta_j p_{j-1},   w = B^-1 r,   beta_{j+1}^2 = w^H r,   v_{j+1} = w / beta_{j+1},   p_{j+1} = r / beta_{j+1}
# with alpha_j = Re(v_j^H A v_j).  Pass 2 replays the recurrence (deterministic inner solves) and accumulates Q += c_j v_j.
# ======================================================================================================================
def jacobi_pcg(B, R, tol=1e-14, maxiter=200, stats=None):
    """Lock-step block CG with Jacobi preconditioning for B X = R (B Hermitian positive definite), every column to
    ||r|| <= tol ||r0||; deterministic (pass 2 of the filter replays it)."""
    dinv = 1.0 / np.real(B.diagonal())
    X = np.zeros_like(R)
    r = R.copy()
    z = r * dinv[:, None]
    p = z.copy()
    rz = np.einsum("ij,ij->j", r.conj(), z).real
    r0 = np.linalg.norm(r, axis=0)
    it = 0
    for it in range(1, maxiter + 1):
        Bp = B @ p
        pBp = np.einsum("ij,ij->j", p.conj(), Bp).real
        a = np.where(pBp > 0, rz / np.where(pBp > 0, pBp, 1.0), 0.0)
        X += a * p
        r -= a * Bp
        if np.all(np.linalg.norm(r, axis=0) <= tol * np.maximum(r0, 1e-300)):
            break
        z = r * dinv[:, None]
        rz_new = np.einsum("ij,ij->j", r.conj(), z).real
        p = z + np.where(rz > 0, rz_new / np.where(rz > 0, rz, 1.0), 0.0) * p
        rz = rz_new
    if stats is not None:
        stats["pcg_iters"].append(it)
    return X


def mslanczos_filter_gen(A, B, Q, theta, Zne, Wne, target, kmax, stats=None, inner_tol=1e-14):
    """sum_e Herm_B-part of 2 w_e (z_e B - A)^-1 B q for the block Q (complex Hermitian pencil, B HPD); theta = Ritz values or None."""
    n, m = Q.shape
    ne = len(Zne)
    solveB = lambda R: jacobi_pcg(B, R, inner_tol, stats=stats)
    if theta is None:
        b, F, acc = Q.astype(complex), np.ones((ne, m), dtype=complex), np.zeros((n, m), dtype=complex)
    else:
        b = solveB(A @ Q - (B @ Q) * theta)                               # C q - theta q
        F = 1.0 / (np.asarray(Zne)[:, None] - theta[None, :])
        acc = Q * np.real((2 * np.asarray(Wne)[:, None] * F).sum(axis=0))
    pb = B @ b
    beta0 = np.sqrt(np.maximum(np.einsum("ij,ij->j", b.conj(), pb).real, 0.0))
    inv0 = np.where(beta0 > 1e-290, 1.0 / np.where(beta0 > 0, beta0, 1.0), 0.0)

    def run(k_fixed=None, alpha=None, beta=None, coef=None, out=None):
        al_l, be_l = [], [beta0]
        v, p = b * inv0, pb * inv0
        p_prev = np.zeros_like(p)
        bj = np.zeros(m)
        d = np.zeros((ne, m), dtype=complex)
        g = np.zeros((ne, m), dtype=complex)
        scale = np.zeros(m)
        k = 0
        steps = kmax if k_fixed is None else k_fixed
        for j in range(steps):
            if out is not None:
                out += coef[j] * v
                if j == steps - 1:
                    break
            Av = A @ v
            if alpha is None:
                al = np.einsum("ij,ij->j", v.conj(), Av).real
            else:
                al = alpha[j]
            r = Av - al * p - bj * p_prev
            w = solveB(r)
            if beta is None:
                bn2 = np.einsum("ij,ij->j", w.conj(), r).real
                bn = np.sqrt(np.maximum(bn2, 0.0))
                scale = np.maximum(scale, np.maximum(np.abs(al), bn))
                ok = (bn > 1e-290) & (bn > 1e-13 * scale)
                bn = np.where(ok, bn, 0.0)
            else:
                bn = beta[j + 1]
            invn = np.where(bn > 0, 1.0 / np.where(bn > 0, bn, 1.0), 0.0)
            if alpha is None:
                al_l.append(al)
                be_l.append(bn)
                for e in range(ne):
                    if j == 0:
                        d[e] = Zne[e] - al
                        g[e] = 1.0 / d[e]
                    else:
                        dn = (Zne[e] - al) - bj ** 2 / d[e]
                        g[e] = bj * g[e] / dn
                        d[e] = dn
                k = j + 1
                if float((bn[None, :] * np.abs(g)).max()) <= target:
                    break
            p_prev, p, v, bj = p, r * invn, w * invn, bn
        return np.array(al_l), np.array(be_l), k

    alpha, beta, k = run()
    beta_T = beta.copy()
    beta_T[0] = beta0                                                        # lanczos_coefficients scales by ||b|| = beta[0]
    coef = lanczos_coefficients(alpha, beta_T[:k + 1], Zne, Wne, F)
    # lanczos_coefficients returns c_j / beta_j for UNNORMALISED vectors u_j = beta_j v_j; here the vectors are normalised
    coef = coef * beta_T[:k]
    run(k_fixed=k, alpha=alpha, beta=beta, coef=coef, out=acc)
    if stats is not None:
        stats["lz_steps"].append(k)
    return acc


def feast_hrr_mslanczos_gen(A, B, Emin, Emax, M0, fpm, Q0, inner_rel=1e-3, inner_maxiter=2000, adaptive=True, verbose=False,
                            b_solve_tol=1e-8):
    """H-RR refinement loop for the generalized Hermitian problem with the B-inner-product multi-shift Lanczos filter.
    b_solve_tol: relative accuracy of the inner solves with B (the recurrence tolerates 1e-6: measured 2 sweeps at every setting
    between 1e-14 and 1e-6 on the reduced configs[3] pair, a few more Lanczos steps for the loosest)."""
    import scipy.linalg as sla
    N = A.shape[0]
    fo.feastdefault(fpm)
    fo.check_feast_srci_input(N, M0, Emin, Emax, fpm)
    Zne, Wne = fo.feast_contour(Emin, Emax, fpm)
    eps_tol = fo.feast_tolerance(fpm)
    Qb = np.array(Q0, dtype=complex)
    lam, res = np.zeros(M0), np.zeros(M0)
    X = np.zeros((N, M0), dtype=complex)
    have_ritz, active = False, M0
    info, epsout, loop_count, M_found = fo.SUCCESS, math.inf, 0, 0
    stats = {"lz_steps": [], "pcg_iters": []}
    for loop_idx in range(fpm[3] + 1):
        loop_count = loop_idx
        target = inner_rel
        if adaptive and have_ritz and math.isfinite(epsout) and epsout > 0:
            t = 2.0 * eps_tol / epsout
            if t >= 1e-6:
                target = min(0.1, t)
        acc = mslanczos_filter_gen(A, B, Qb[:, :active], lam[:active].copy() if have_ritz else None, Zne, Wne, target, inner_maxiter, stats,
                                   b_solve_tol)
        Qr, rank = fo.qr_compress(np.ascontiguousarray(acc), active)
        if rank == 0:
            info = fo.ERR_NO_CONV
            break
        Aq = Qr.conj().T @ (A @ Qr)
        Bq = Qr.conj().T @ (B @ Qr)
        lam_red, v_red = sla.eigh(0.5 * (Aq + Aq.conj().T), 0.5 * (Bq + Bq.conj().T))
        Xc = np.zeros((N, M0), dtype=complex)
        Xc[:, :rank] = Qr @ v_red
        lam[:rank] = lam_red
        M = fo.reorder_by_interval(lam, Xc, Emin, Emax, rank)
        X = Xc
        if M == 0:
            info = fo.ERR_NO_CONV
            break
        X[:, :M] /= np.linalg.norm(X[:, :M], axis=0)
        R = A @ X[:, :M] - (B @ X[:, :M]) * lam[:M]
        res[:M] = np.linalg.norm(R, axis=0) / np.maximum(np.abs(lam[:M]), 1.0)
        epsout = float(res[:M].max())
        M_found = M
        if verbose:
            print(f"loop {loop_idx}: M={M} rank={rank} epsout={epsout:.3e} k={stats['lz_steps'][-1]} pcg~{int(np.mean(stats['pcg_iters'][-5:]))}", flush=True)
        if epsout <= eps_tol:
            break
        if loop_idx == fpm[3]:
            info = fo.ERR_NO_CONV
            break
        active = rank
        Qb = X[:, :active].copy()
        have_ritz = True
    return fo.FeastResult(lam[:M_found].copy(), X[:, :M_found].copy(), M_found, res[:M_found].copy(), info, epsout, loop_count, stats)


# ======================================================================================================================
# General (non-Hermitian) pencils with a SHARED Krylov space -- the design for BASELINE configs[4],
# restated on the CPU before any CUDA exists for it.  (z B - A)^-1 B q = (z I - C)^-1 q with C = B^-1 A, and the Arnoldi
# relation C V_k = V_k H_k + h_{k+1,k} v_{k+1} e_k^T does not depend on z: one Arnoldi run per column (lock step, modified
# Gram-Schmidt) serves every quadrature node through the small shifted Hessenberg solves (multi-shift FOM)
#     x_e = ||b|| V_k (z_e I - H_k)^-1 e_1,      residual_e = h_{k+1,k} |e_k^T (z_e I - H_k)^-1 e_1| ||b||.
# The basis V_k is STORED (k x n x m complex: 25 GB for k = 100 at configs[4] -- what 180 GB of HBM are for), so there is no
# second pass; the per-node Krylov solves the reference runs (ne x m GMRES calls) collapse into one.
# ======================================================================================================================
def msarnoldi_filter(A, solveB, Q, theta, Zne, Wne, target, kmax, stats=None, check_every=4):
    """sum_e w_e (z_e B - A)^-1 B q for the block Q (full contour, complex weights); theta = Ritz values (residual start) or None."""
    n, m = Q.shape
    ne = len(Zne)
    Z, W = np.asarray(Zne), np.asarray(Wne)
    applyC = lambda X: solveB(A @ X)
    if theta is None:
        b, F, acc = Q.astype(complex), np.ones((ne, m), dtype=complex), np.zeros((n, m), dtype=complex)
    else:
        b = applyC(Q) - Q * theta
        F = 1.0 / (Z[:, None] - theta[None, :])
        acc = Q * (W[:, None] * F).sum(axis=0)
    beta0 = np.linalg.norm(b, axis=0)
    V = [b / np.where(beta0 > 0, beta0, 1.0)]
    H = np.zeros((kmax + 1, kmax, m), dtype=complex)
    k = 0
    for j in range(kmax):
        w = applyC(V[j])
        for i in range(j + 1):                              # modified Gram-Schmidt, every column on its own basis
            hij = np.einsum("ij,ij->j", V[i].conj(), w)
            H[i, j] = hij
            w = w - hij * V[i]
        hn = np.linalg.norm(w, axis=0)
        H[j + 1, j] = hn
        k = j + 1
        # shifted FOM residuals of every (node, column), looked at every few steps (k^3 work per pair on the host)
        breakdown = not np.all(hn > 1e-13 * np.abs(H[:k, :k]).max())
        if k % check_every == 0 or k == kmax or breakdown:
            worst = 0.0
            e1 = np.eye(k)[:, 0]
            for c in range(m):
                ys = np.linalg.solve(Z[:, None, None] * np.eye(k)[None] - H[:k, :k, c][None], np.broadcast_to(e1, (ne, k))[..., None])[:, -1, 0]
                worst = max(worst, float(hn[c] * np.abs(ys).max()))
            if worst <= target or breakdown:
                break
        V.append(w / np.where(hn > 0, hn, 1.0))
    for c in range(m):
        if not beta0[c] > 0:
            continue
        Hk = H[:k, :k, c]
        coef = np.zeros(k, dtype=complex)
        for e in range(ne):
            coef += W[e] * F[e, c] * np.linalg.solve(Z[e] * np.eye(k) - Hk, np.eye(k)[:, 0])
        coef *= beta0[c]
        for j in range(k):
            acc[:, c] += coef[j] * V[j][:, c]
    if stats is not None:
        stats["arnoldi_steps"].append(k)
    return acc


def feast_general_msarnoldi(A, B, Emid, r, M0, fpm, Q0, inner_rel=1e-3, inner_maxiter=200, adaptive=True, verbose=False):
    """General FEAST (full contour, one-sided Rayleigh-Ritz on the ORTHONORMALISED filtered block as the engine's run_contour does,
    residuals with B) with the multi-shift Arnoldi filter; B^-1 by sparse LU here (the engine would use a few Richardson or
    Krylov steps: configs[4]'s B = I + 0.05 S is a small perturbation of the identity)."""
    import scipy.linalg as sla
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    N = A.shape[0]
    fo.feastdefault(fpm)
    fo.check_feast_grci_input(N, M0, Emid, r, fpm)
    Emid = complex(Emid)
    Zne, Wne = fo.feast_gcontour(Emid, r, fpm)
    eps_tol = fo.feast_tolerance(fpm)
    Ac = sp.csr_matrix(A, dtype=complex)
    Bc = sp.identity(N, dtype=complex, format="csr") if B is None else sp.csr_matrix(B, dtype=complex)
    luB = spla.splu(Bc.tocsc())
    solveB = (lambda X: X) if B is None else (lambda X: luB.solve(np.ascontiguousarray(X)))
    Qb = np.array(Q0, dtype=complex)
    lam, res = np.zeros(M0, dtype=complex), np.zeros(M0)
    X = np.zeros((N, M0), dtype=complex)
    have_ritz, active = False, M0
    info, epsout, loop_count, M_found = fo.SUCCESS, math.inf, 0, 0
    stats = {"arnoldi_steps": []}
    for loop_idx in range(fpm[3] + 1):
        loop_count = loop_idx
        target = inner_rel
        if adaptive and have_ritz and math.isfinite(epsout) and epsout > 0:
            t = 2.0 * eps_tol / epsout
            if t >= 1e-6:
                target = min(0.1, t)
        acc = msarnoldi_filter(Ac, solveB, Qb[:, :active], lam[:active].copy() if have_ritz else None, Zne, Wne, target, inner_maxiter,
                               stats)
        Qr, rank = fo.qr_compress(np.ascontiguousarray(acc), active)
        if rank == 0:
            info = fo.ERR_NO_CONV
            break
        Aq = Qr.conj().T @ (Ac @ Qr)
        Bq = Qr.conj().T @ (Bc @ Qr)
        lam_red, v_red = sla.eig(Aq, Bq)
        Xc = np.zeros((N, M0), dtype=complex)
        Xc[:, :rank] = Qr @ v_red
        lam[:rank] = lam_red
        M = fo.reorder_by_gcontour(lam, Xc, Emid, r, fpm, rank)
        if M == 0:
            info = fo.ERR_NO_CONV
            break
        nrm = np.linalg.norm(Xc[:, :rank], axis=0)
        Xc[:, :rank] /= np.where(nrm > 0, nrm, 1.0)
        X = Xc
        R = Ac @ X[:, :M] - (Bc @ X[:, :M]) * lam[:M]
        res[:M] = np.linalg.norm(R, axis=0) / np.maximum(np.abs(lam[:M]), 1.0)
        epsout = float(res[:M].max())
        M_found = M
        if verbose:
            print(f"loop {loop_idx}: M={M} rank={rank} epsout={epsout:.3e} k={stats['arnoldi_steps'][-1]}", flush=True)
        if epsout <= eps_tol:
            break
        if loop_idx == fpm[3]:
            info = fo.ERR_NO_CONV
            break
        active = rank
        Qb = X[:, :active].copy()
        have_ritz = True
    lam_out, q_out, res_out = lam[:M_found].copy(), X[:, :M_found].copy(), res[:M_found].copy()
    fo.feast_sort_general(lam_out, q_out, res_out, M_found)
    return fo.FeastResult(lam_out, q_out, M_found, res_out, info, epsout, loop_count, stats)


# ======================================================================================================================
# The generalized Hermitian filter as the CUDA engine runs it (csrc/feastcuda.cu:msl_filter_gen): the inner solves with B are
# a FIXED Chebyshev polynomial P = p_K(D^-1 B) D^-1 (D = diag B; no dot products, no data-dependent branches), so the recurrence
# is an EXACT Lanczos process for the neighbouring pencil (A, P^-1): self-adjoint in the P^-1 inner product, which is never
# applied -- the images s_j = P^-1 u_j ride along (s_{j+1} is the right-hand side of the solve that produced u_{j+1}).
# Vectors are unnormalised (u_j = beta_j v_j, s_j = beta_j P^-1 v_j) like in the standard path, the scalars are the same
# (alpha_j beta_j = u_j^H t, beta_{j+1}^2 = u_{j+1}^H s_{j+1}), and so are the host-side tridiagonal solves.
# ======================================================================================================================
def chebyshev_setup(B, steps=40, safety=0.85):
    """Spectral interval [lo, hi] of D^-1 B for the Chebyshev inner solver: hi from Gershgorin's bound (safe), lo from the smallest Ritz
    value of a short Lanczos run on D^-1/2 B D^-1/2 times a safety factor (a Ritz value bounds the smallest eigenvalue from above)."""
    import scipy.sparse as sp
    import scipy.linalg as sla
    d = np.real(B.diagonal())
    assert np.all(d > 0)
    dinv = 1.0 / d
    absB = abs(sp.csr_matrix(B))
    hi = float((absB @ np.ones(B.shape[0]) * dinv).max())
    n = B.shape[0]
    s = np.sqrt(dinv)
    v = np.cos(1.0 + np.arange(n) * 0.7548776662466927)     # deterministic quasi-random start (golden-ratio-like increment)
    v = v / np.linalg.norm(v)
    v_prev = np.zeros(n)
    al, be = [], []
    beta = 0.0
    for _ in range(min(steps, n)):
        w = s * (B @ (s * v))
        w = np.real(w) if not np.iscomplexobj(B) else w
        a = float(np.real(np.vdot(v, w)))
        w = w - a * v - beta * v_prev
        beta = float(np.linalg.norm(w))
        al.append(a)
        if beta < 1e-12 * hi:
            break
        be.append(beta)
        v_prev, v = v, w / beta
    k = len(al)
    ritz = sla.eigvalsh_tridiagonal(np.array(al), np.array(be[:k - 1])) if k > 1 else np.array(al)
    lo = max(safety * float(ritz.min()), 1e-3 * hi)
    return dinv, lo, hi


def chebyshev_degree(lo, hi, delta):
    """Smallest K with the Chebyshev error bound 2 s^K / (1 + s^2K) <= delta, s = (sqrt(kappa)-1)/(sqrt(kappa)+1)."""
    kap = hi / lo
    s = (math.sqrt(kap) - 1.0) / (math.sqrt(kap) + 1.0)
    K = 1
    while 2.0 * s ** K / (1.0 + s ** (2 * K)) > delta and K < 400:
        K += 1
    return K


def chebyshev_solve(B, dinv, lo, hi, R, K):
    """X = p_K(D^-1 B) D^-1 R: K steps of the three-term Chebyshev iteration from X_0 = 0 (k_lz_spmm<LZ_CHEB>)."""
    theta, delta = 0.5 * (hi + lo), 0.5 * (hi - lo)
    sigma = theta / delta
    x_prev = np.zeros_like(R)
    x = (dinv[:, None] * R) / theta
    rho_prev = 1.0 / sigma
    for _ in range(1, K):
        rho = 1.0 / (2.0 * sigma - rho_prev)
        x_new = x + rho * rho_prev * (x - x_prev) + (2.0 * rho / delta) * (dinv[:, None] * (R - B @ x))
        x_prev, x, rho_prev = x, x_new, rho
    return x


def mslanczos_filter_gen_cheb(A, B, cheb, K, Q, theta, Zne, Wne, target, kmax, stats=None):
    """Engine-order restatement: start s_0 = B q (or A q - theta B q), u_0 = P s_0; step: t = A u_j / beta_j - (beta_j/beta_{j-1}) s_{j-1},
    alpha_j beta_j = Re(u_j^H t), s_{j+1} = t - (alpha_j/beta_j) s_j, u_{j+1} = P s_{j+1}, beta_{j+1}^2 = Re(u_{j+1}^H s_{j+1})."""
    dinv, lo, hi = cheb
    n, m = Q.shape
    ne = len(Zne)
    P = lambda R: chebyshev_solve(B, dinv, lo, hi, R, K)
    Qc = Q.astype(complex)
    if theta is None:
        s0, F, acc = B @ Qc, np.ones((ne, m), dtype=complex), np.zeros((n, m), dtype=complex)
    else:
        s0 = A @ Qc - (B @ Qc) * theta
        F = 1.0 / (np.asarray(Zne)[:, None] - theta[None, :])
        acc = Qc * np.real((2 * np.asarray(Wne)[:, None] * F).sum(axis=0))
    u0 = P(s0)
    beta0 = np.sqrt(np.maximum(np.einsum("ij,ij->j", u0.conj(), s0).real, 0.0))

    def run(alpha=None, beta=None, coef=None, out=None):
        al_l, be_l = [], [beta0]
        u, s, s_prev = u0, s0, np.zeros_like(s0)
        bj = beta0
        inv = np.where(bj > 1e-290, 1.0 / np.where(bj > 0, bj, 1.0), 0.0)
        ratio_b = np.zeros(m)
        scale = np.zeros(m)
        d = np.zeros((ne, m), dtype=complex)
        g = np.zeros((ne, m), dtype=complex)
        k = 0
        steps = kmax if alpha is None else len(alpha)
        for j in range(steps):
            if out is not None:
                out += coef[j] * u
                if j == steps - 1:
                    break
            t = (A @ u) * inv - ratio_b * s_prev
            if alpha is None:
                al = np.einsum("ij,ij->j", u.conj(), t).real * inv
                al_l.append(al)
                scale = np.maximum(scale, np.abs(al))
            else:
                al = alpha[j]
            s_new = t - (al * inv) * s
            u_new = P(s_new)
            if alpha is None:
                bn = np.sqrt(np.maximum(np.einsum("ij,ij->j", u_new.conj(), s_new).real, 0.0))
                ok = (bn > 1e-290) & (bn > 1e-13 * scale) & (inv != 0.0)
                bn = np.where(ok, bn, 0.0)
                be_l.append(bn)
                scale = np.maximum(scale, bn)
                for e in range(ne):
                    if j == 0:
                        d[e] = Zne[e] - al
                        g[e] = 1.0 / d[e]
                    else:
                        dn = (Zne[e] - al) - bj ** 2 / d[e]
                        g[e] = bj * g[e] / dn
                        d[e] = dn
                k = j + 1
            else:
                bn = beta[j + 1]
            inv_new = np.where(bn > 0, 1.0 / np.where(bn > 0, bn, 1.0), 0.0)
            ratio_b = np.where(bn > 0, bn * inv, 0.0)
            s_prev, s, u, bj, inv = s, s_new, u_new, bn, inv_new
            if alpha is None and float((bn[None, :] * np.abs(g)).max()) <= target:
                break
        return np.array(al_l), np.array(be_l), k

    alpha, beta, k = run()
    coef = lanczos_coefficients(alpha, beta[:k + 1], Zne, Wne, F)          # c_j / beta_j for the unnormalised u_j
    run(alpha=alpha, beta=beta, coef=coef, out=acc)
    if stats is not None:
        stats["lz_steps"].append(k)
        stats["cheb_degree"].append(K)
    return acc


def feast_hrr_mslanczos_gen_cheb(A, B, Emin, Emax, M0, fpm, Q0, inner_rel=1e-3, inner_maxiter=4000, adaptive=True, verbose=False,
                                 b_delta0=1e-4, b_delta_min=1e-9):
    """H-RR loop around mslanczos_filter_gen_cheb, as run_interval drives msl_filter_gen.  The Chebyshev accuracy follows the outer
    residual: delta = b_delta0 in the first sweep, then clamp(1e-2 * epsout, b_delta_min, b_delta0) -- the filter of the pencil
    (A, P^-1) differs from the exact one by O(delta |z| / dist), which only has to stay below the sweep's contraction."""
    import scipy.linalg as sla
    N = A.shape[0]
    fo.feastdefault(fpm)
    fo.check_feast_srci_input(N, M0, Emin, Emax, fpm)
    Zne, Wne = fo.feast_contour(Emin, Emax, fpm)
    eps_tol = fo.feast_tolerance(fpm)
    cheb = chebyshev_setup(B)
    Qb = np.array(Q0, dtype=complex)
    lam, res = np.zeros(M0), np.zeros(M0)
    X = np.zeros((N, M0), dtype=complex)
    have_ritz, active = False, M0
    info, epsout, loop_count, M_found = fo.SUCCESS, math.inf, 0, 0
    stats = {"lz_steps": [], "cheb_degree": [], "cheb_interval": cheb[1:]}
    for loop_idx in range(fpm[3] + 1):
        loop_count = loop_idx
        target, delta = inner_rel, b_delta0
        if have_ritz and math.isfinite(epsout) and epsout > 0:
            delta = min(b_delta0, max(b_delta_min, 1e-2 * epsout))
            if adaptive:
                t = 2.0 * eps_tol / epsout
                if t >= 1e-6:
                    target = min(0.1, t)
        K = chebyshev_degree(cheb[1], cheb[2], delta)
        acc = mslanczos_filter_gen_cheb(A, B, cheb, K, Qb[:, :active], lam[:active].copy() if have_ritz else None, Zne, Wne, target,
                                        inner_maxiter, stats)
        Qr, rank = fo.qr_compress(np.ascontiguousarray(acc), active)
        if rank == 0:
            info = fo.ERR_NO_CONV
            break
        Aq = Qr.conj().T @ (A @ Qr)
        Bq = Qr.conj().T @ (B @ Qr)
        lam_red, v_red = sla.eigh(0.5 * (Aq + Aq.conj().T), 0.5 * (Bq + Bq.conj().T))
        Xc = np.zeros((N, M0), dtype=complex)
        Xc[:, :rank] = Qr @ v_red
        lam[:rank] = lam_red
        M = fo.reorder_by_interval(lam, Xc, Emin, Emax, rank)
        X = Xc
        if M == 0:
            info = fo.ERR_NO_CONV
            break
        X[:, :M] /= np.linalg.norm(X[:, :M], axis=0)
        R = A @ X[:, :M] - (B @ X[:, :M]) * lam[:M]
        res[:M] = np.linalg.norm(R, axis=0) / np.maximum(np.abs(lam[:M]), 1.0)
        epsout = float(res[:M].max())
        M_found = M
        if verbose:
            print(f"loop {loop_idx}: M={M} rank={rank} epsout={epsout:.3e} k={stats['lz_steps'][-1]} K={K} delta={delta:.1e}", flush=True)
        if epsout <= eps_tol:
            break
        if loop_idx == fpm[3]:
            info = fo.ERR_NO_CONV
            break
        active = rank
        Qb = X[:, :active].copy()
        have_ritz = True
    return fo.FeastResult(lam[:M_found].copy(), X[:, :M_found].copy(), M_found, res[:M_found].copy(), info, epsout, loop_count, stats)


# ======================================================================================================================
# General pencils as the CUDA engine runs them (csrc/feastcuda.cu:msl2_filter): TWO-SIDED multi-shift Lanczos, two passes, no stored
# basis.  C = B'^-1 A' with the rows scaled by diag(B)^-1 and B'^-1 a FIXED number of Jacobi sweeps (a polynomial in B', so the
# adjoint operator of the w-sequence is the same polynomial in B'^H).  ||v_j|| = ||w_j|| = 1, d_j = w_j^H v_j, T_k complex tridiagonal:
#     C v_j = beta'_j v_{j-1} + alpha_j v_j + delta_{j+1} v_{j+1},      beta'_{j+1} = gamma_{j+1} d_{j+1} / d_j,
#     C^H w_j = conj(delta_j d_j / d_{j-1}) w_{j-1} + conj(alpha_j) w_j + gamma_{j+1} w_{j+1}.
# ======================================================================================================================
def jacobi_setup(A, B):
    """Row-scaled operators (A', B') with unit diagonal of B', their conjugate transposes, and the contraction bound q of I - B'."""
    import scipy.sparse as sp
    A = sp.csr_matrix(A, dtype=complex)
    n = A.shape[0]
    if B is None:
        return A, None, A.conj().T.tocsr(), None, 0.0
    B = sp.csr_matrix(B, dtype=complex)
    d = B.diagonal()
    assert np.all(np.abs(d) > 0)
    Dinv = sp.diags(1.0 / d)
    Ap, Bp = (Dinv @ A).tocsr(), (Dinv @ B).tocsr()
    off = abs(Bp - sp.identity(n, dtype=complex, format="csr"))
    q = max(float(off.sum(axis=1).max()), float(off.sum(axis=0).max()))
    return Ap, Bp, Ap.conj().T.tocsr(), Bp.conj().T.tocsr(), q


def jacobi_solve(M, R, K):
    """K Jacobi sweeps from x_1 = r for a unit-diagonal M (k_lz_spmm<LZ_CHEB> with c1 = 0, c2 = 1)."""
    x = R
    for _ in range(1, K):
        x = x + (R - M @ x)
    return x


def mstwosided_filter(ops, Q, theta, Zne, Wne, target, kmax, jac_delta=1e-10, stats=None):
    """sum_e w_e (z_e B - A)^-1 B q for the block Q (full contour); theta = Ritz values (residual start) or None."""
    Ap, Bp, Ah, Bh, q = ops
    n, m = Q.shape
    ne = len(Zne)
    Z, W = np.asarray(Zne), np.asarray(Wne)
    KJ = 1 if Bp is None or q <= 0 else max(1, int(math.ceil(math.log(jac_delta) / math.log(q))))
    applyC = (lambda X: Ap @ X) if Bp is None else (lambda X: jacobi_solve(Bp, Ap @ X, KJ))
    applyCh = (lambda X: Ah @ X) if Bp is None else (lambda X: Ah @ jacobi_solve(Bh, X, KJ))
    Qc = Q.astype(complex)
    if theta is None:
        b, F, acc = Qc, np.ones((ne, m), dtype=complex), np.zeros((n, m), dtype=complex)
    else:
        b = applyC(Qc) - Qc * theta
        F = 1.0 / (Z[:, None] - theta[None, :])
        acc = Qc * (W[:, None] * F).sum(axis=0)
    beta0 = np.linalg.norm(b, axis=0)
    v0 = b / np.where(beta0 > 0, beta0, 1.0)
    alpha = np.zeros((kmax + 2, m), dtype=complex)
    betap = np.zeros((kmax + 2, m), dtype=complex)
    delta = np.zeros((kmax + 2, m))
    dd = np.ones((kmax + 2, m), dtype=complex)
    kc = np.zeros(m, dtype=int)
    v, w = v0.copy(), v0.copy()
    vp, wp = np.zeros_like(v), np.zeros_like(w)
    lud = np.zeros((ne, m), dtype=complex)
    lug = np.zeros((ne, m), dtype=complex)
    scale = np.zeros(m)
    k = 0
    for j in range(kmax):
        alive = (kc == 0) & (beta0 > 0)
        cv = applyC(v)
        al = np.where(alive, np.einsum("ij,ij->j", w.conj(), cv) / dd[j], 0.0)
        alpha[j] = al
        scale = np.maximum(scale, np.abs(al))
        vh = cv - al * v - betap[j] * vp
        cw = applyCh(w)
        gmc = np.conj(delta[j] * dd[j] / dd[j - 1]) if j > 0 else np.zeros(m)
        wh = cw - np.conj(al) * w - np.where(alive, gmc, 0.0) * wp
        dl, gm = np.linalg.norm(vh, axis=0), np.linalg.norm(wh, axis=0)
        ok = alive & (dl > 1e-290) & (gm > 1e-290) & (dl > 1e-13 * np.maximum(scale, 1e-300))
        dn = np.where(ok, np.einsum("ij,ij->j", wh.conj(), vh) / np.where(ok, dl * gm, 1.0), 1.0)
        stop = alive & (~ok | ~(np.abs(dn) > 1e-10))
        kc = np.where(stop, j + 1, kc)
        dl = np.where(alive & ~ok, 0.0, dl)
        live = alive & ~stop
        scale = np.maximum(scale, np.where(alive, dl, 0.0))
        delta[j + 1] = np.where(alive, dl, 0.0)
        dd[j + 1] = np.where(live, dn, 1.0)
        betap[j + 1] = np.where(live, gm * dn / dd[j], 0.0)
        worst = 0.0
        for e in range(ne):
            if j == 0:
                d = Z[e] - al
                g = 1.0 / d
            else:
                d = (Z[e] - al) - betap[j] * delta[j] / lud[e]
                g = delta[j] * lug[e] / d
            upd = alive
            lud[e] = np.where(upd, d, lud[e])
            lug[e] = np.where(upd, g, lug[e])
            worst = max(worst, float(np.where(upd, delta[j + 1] * np.abs(g), 0.0).max()))
        vp, wp = v, w
        v = np.where(live, vh / np.where(live, dl, 1.0), 0.0)
        w = np.where(live, wh / np.where(live, gm, 1.0), 0.0)
        k = j + 1
        if not worst > target:
            break
    coef = np.zeros((k, m), dtype=complex)
    for c in range(m):
        if not beta0[c] > 0:
            continue
        kk = min(kc[c], k) if kc[c] else k
        for e in range(ne):
            wf = W[e] * F[e, c] * beta0[c]
            ludv = np.zeros(kk, dtype=complex)
            luf = np.zeros(kk, dtype=complex)
            ludv[0], luf[0] = Z[e] - alpha[0, c], 1.0
            for j in range(1, kk):
                l = delta[j, c] / ludv[j - 1]
                ludv[j] = (Z[e] - alpha[j, c]) - l * betap[j, c]
                luf[j] = l * luf[j - 1]
            y = luf[kk - 1] / ludv[kk - 1]
            coef[kk - 1, c] += wf * y
            for j in range(kk - 2, -1, -1):
                y = (luf[j] + betap[j + 1, c] * y) / ludv[j]
                coef[j, c] += wf * y
    # pass 2: the v-sequence alone
    v, vp = v0.copy(), np.zeros_like(v0)
    for j in range(k):
        acc += coef[j] * v
        if j == k - 1:
            break
        vh = applyC(v) - alpha[j] * v - betap[j] * vp
        live = (beta0 > 0) & ((kc == 0) | (j + 1 < kc)) & (delta[j + 1] > 0)
        vp, v = v, np.where(live, vh / np.where(live, delta[j + 1], 1.0), 0.0)
    if stats is not None:
        stats["lz_steps"].append(k)
        stats["jacobi_sweeps"] = KJ
    return acc


def feast_general_mstwosided(A, B, Emid, r, M0, fpm, Q0, inner_rel=1e-3, inner_maxiter=2000, adaptive=True, verbose=False, jac_delta=1e-10):
    """General FEAST (run_contour's loop: orthonormalised block, one-sided Rayleigh-Ritz, residuals with B) around mstwosided_filter."""
    import scipy.linalg as sla
    import scipy.sparse as sp
    N = A.shape[0]
    fo.feastdefault(fpm)
    fo.check_feast_grci_input(N, M0, Emid, r, fpm)
    Emid = complex(Emid)
    Zne, Wne = fo.feast_gcontour(Emid, r, fpm)
    eps_tol = fo.feast_tolerance(fpm)
    Ac = sp.csr_matrix(A, dtype=complex)
    Bc = sp.identity(N, dtype=complex, format="csr") if B is None else sp.csr_matrix(B, dtype=complex)
    ops = jacobi_setup(A, B)
    Qb = np.array(Q0, dtype=complex)
    lam, res = np.zeros(M0, dtype=complex), np.zeros(M0)
    X = np.zeros((N, M0), dtype=complex)
    have_ritz, active = False, M0
    info, epsout, loop_count, M_found = fo.SUCCESS, math.inf, 0, 0
    stats = {"lz_steps": []}
    for loop_idx in range(fpm[3] + 1):
        loop_count = loop_idx
        target = inner_rel
        if adaptive and have_ritz and math.isfinite(epsout) and epsout > 0:
            t = 2.0 * eps_tol / epsout
            if t >= 1e-6:
                target = min(0.1, t)
        acc = mstwosided_filter(ops, Qb[:, :active], lam[:active].copy() if have_ritz else None, Zne, Wne, target, inner_maxiter, jac_delta, stats)
        Qr, rank = fo.qr_compress(np.ascontiguousarray(acc), active)
        if rank == 0:
            info = fo.ERR_NO_CONV
            break
        lam_red, v_red = sla.eig(Qr.conj().T @ (Ac @ Qr), Qr.conj().T @ (Bc @ Qr))
        Xc = np.zeros((N, M0), dtype=complex)
        Xc[:, :rank] = Qr @ v_red
        lam[:rank] = lam_red
        M = fo.reorder_by_gcontour(lam, Xc, Emid, r, fpm, rank)
        if M == 0:
            info = fo.ERR_NO_CONV
            break
        nrm = np.linalg.norm(Xc[:, :rank], axis=0)
        Xc[:, :rank] /= np.where(nrm > 0, nrm, 1.0)
        X = Xc
        R = Ac @ X[:, :M] - (Bc @ X[:, :M]) * lam[:M]
        res[:M] = np.linalg.norm(R, axis=0) / np.maximum(np.abs(lam[:M]), 1.0)
        epsout = float(res[:M].max())
        M_found = M
        if verbose:
            print(f"loop {loop_idx}: M={M} rank={rank} epsout={epsout:.3e} k={stats['lz_steps'][-1]} KJ={stats.get('jacobi_sweeps')}", flush=True)
        if epsout <= eps_tol:
            break
        if loop_idx == fpm[3]:
            info = fo.ERR_NO_CONV
            break
        active = rank
        Qb = X[:, :active].copy()
        have_ritz = True
    lam_out, q_out, res_out = lam[:M_found].copy(), X[:, :M_found].copy(), res[:M_found].copy()
    fo.feast_sort_general(lam_out, q_out, res_out, M_found)
    return fo.FeastResult(lam_out, q_out, M_found, res_out, info, epsout, loop_count, stats)


# =====================================================================================================================
# banded direct solves: port of csrc/kernels_band.cuh (k_band_shift, k_band_lu_warp, k_band_solve_win)
# =====================================================================================================================
def band_shift(AB, BB, z, ka, kb):
    """F = z B - A in the factor layout of the engine: (3k+1) x n, k = max(ka, kb), entry (i, j) at F[2k + i - j, j], the top k rows zero
    (fill-in of the pivoted factorisation); rows that fall outside the matrix are zero.  AB / BB: general band (2k+1) x n, diagonal in
    row k (BB None: B = I).  Replaces fill_shifted_banded! (banded/feast_banded.jl:216-237, 273-296, 511-559)."""
    n = AB.shape[1]
    k = max(ka, kb if BB is not None else 0)
    F = np.zeros((3 * k + 1, n), dtype=complex)
    for j in range(n):
        for i in range(max(0, j - k), min(n - 1, j + k) + 1):
            d = i - j
            v = 0j
            if -ka <= d <= ka:
                v -= AB[ka + d, j]
            if BB is not None:
                if -kb <= d <= kb:
                    v += z * BB[kb + d, j]
            elif d == 0:
                v += z
            F[2 * k + d, j] = v
    return F, k


def band_lu(F, k):
    """Unblocked band LU with partial pivoting (LAPACK zgbtf2, kl = ku = k) in place; returns (ipiv 0-based, info).  The CUDA kernel runs
    exactly these steps with one warp per quadrature node: pivot = first entry of largest |re| + |im| in the column, row swap over the
    columns j..ju, scaling, rank-1 update of the (km x (ju - j)) window."""
    ldf, n = F.shape
    kv = 2 * k
    ipiv = np.zeros(n, dtype=np.int64)
    ju, info = 0, 0
    for j in range(n):
        km = min(k, n - 1 - j)
        col = F[kv:kv + km + 1, j]
        mag = np.abs(col.real) + np.abs(col.imag)
        jp = int(np.argmax(mag))                      # first maximum
        ipiv[j] = j + jp
        ju = max(ju, min(j + k + jp, n - 1))
        if not mag[jp] > 0.0:
            info = info or j + 1
            continue
        if jp:
            for c in range(j, ju + 1):
                r0 = kv + j - c
                F[r0, c], F[r0 + jp, c] = F[r0 + jp, c], F[r0, c]
        F[kv + 1:kv + km + 1, j] /= F[kv, j]
        for c in range(j + 1, ju + 1):
            r0 = kv + j - c                           # row of A[j, c] in column c
            F[r0 + 1:r0 + km + 1, c] -= F[kv + 1:kv + km + 1, j] * F[r0, c]
    return ipiv, info


def band_solve_window(F, ipiv, k, K, rhs):
    """x = (L U)^-1 P rhs for ONE right-hand-side column the way k_band_solve_win<K> does it: the rows a step touches live in a window
    (K + 1 values in the forward sweep, 2K + 1 in the backward sweep, K >= k a compile-time size on the GPU) that slides by one row per
    step -- one load and one store of x per row and sweep; multipliers beyond the bandwidth are skipped by the predicate i <= k, rows
    outside the matrix carry zero factors.  LAPACK zgbtrs (banded/feast_banded.jl:141, 683)."""
    n = F.shape[1]
    kv = 2 * k
    assert K >= k
    y = np.zeros(n, dtype=complex)
    w = np.zeros(K + 1, dtype=complex)
    w[:min(K + 1, n)] = rhs[:min(K + 1, n)]
    for j in range(n):
        p = int(ipiv[j]) - j
        if p:
            w[0], w[p] = w[p], w[0]
        xj = w[0]
        y[j] = xj
        for i in range(1, K + 1):
            if i <= k:
                w[i] -= F[kv + i, j] * xj
        w[:K] = w[1:].copy()
        w[K] = rhs[j + 1 + K] if j + 1 + K < n else 0.0
    x = np.zeros(n, dtype=complex)
    v = np.zeros(2 * K + 1, dtype=complex)
    for d in range(2 * K + 1):
        if n - 1 - d >= 0:
            v[d] = y[n - 1 - d]
    for j in range(n - 1, -1, -1):
        xj = v[0] / F[kv, j]
        x[j] = xj
        for d in range(1, 2 * K + 1):
            if d <= kv:
                v[d] -= F[kv - d, j] * xj
        v[:2 * K] = v[1:].copy()
        v[2 * K] = y[j - 1 - 2 * K] if j - 1 - 2 * K >= 0 else 0.0
    return x
