#!/usr/bin/env python
"""bench.py -- the reference's headline metric on the reference's headline config (BASELINE.json):

    FEAST solve wall-time & eigenpairs/s, sparse 3-D 7-point Laplacian CSR n = 1e6 (100^3), Float64, M0 = 64,
    8 Gauss nodes, interval (0, 0.0222) -> M = 35 eigenpairs, converged to residual < 1e-12.

One "step" = one complete FEAST solve (contour filter by the multi-shift Lanczos inner solver, orthonormalisation,
Rayleigh-Ritz, residual check, refinement loops until epsout <= 10^-fpm[3]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--grid 100] [--m0 64]

N > 1 is launched by the driver under torchrun (one rank per GPU): every rank owns a block of ROWS of A and of every vector
(--shard rows, the default: halo rows pushed into the peers' HBM over NVLink by the producing kernels, dot products as one-shot
peer-memory reductions, Rayleigh-Ritz on local rows; DESIGN.md "Multi-GPU"), or a slice of the RHS columns (--shard columns: ONE
ncclAllReduce of the n x M0 accumulator per refinement loop).  Every rank checks its result (M, eigenvalues against the analytic
spectrum, residuals, and the angle between the computed invariant subspace and the analytic eigenvectors); the line's `result`
carries the worst value over the ranks.

--config 1 | 3 | 4 run the other BASELINE configs at full size on one GPU (dense n = 8192; generalized Hermitian n = 500 000;
general complex n = 250 000) and print a line of the same shape; they are not the headline and take minutes (config 3: ~10 min).

`--impl reference`: the reference is Julia and cannot run in this image (no `julia`), so this arm times the CPU port of
the same solve (oracle/feast_port.py) on the host cores: each step is a bounded sample (S lock-step Lanczos steps of both
passes at full size on all cores), extrapolated to the full solve with the step counts below.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "feastkit.jl_b200"))

METRIC = "FEAST solve eigenpairs/s (sparse 3D Laplacian n=1M, M0=64, 8 nodes; wall-time in ms_per_step)"
UNIT = "eigenpairs/s"
# Lanczos steps the engine needed for this config on B200 (gpurun_out/r1_msl100.log: 718 + 662 + 303 over 3 sweeps, both
# passes each); the CPU arms extrapolate their bounded sample with it.  The live GPU run reports its own count.
C3_LANCZOS_STEPS = 1785
C3_M = 35
SOLVER_KW = dict(solver="mslanczos", inner_rel=1e-3, ritz_guess=True, solver_maxiter=3000, check_every=16, filter="true", adaptive=True)


def laplacian_3d(N):
    import numpy as np
    import scipy.sparse as sp
    T = sp.diags([-np.ones(N - 1), 2.0 * np.ones(N), -np.ones(N - 1)], [-1, 0, 1], format="csr")
    I = sp.identity(N, format="csr")
    return (sp.kron(sp.kron(T, I), I) + sp.kron(sp.kron(I, T), I) + sp.kron(sp.kron(I, I), T)).tocsr()


def laplacian_3d_eigs(N, count):
    import numpy as np
    k = np.arange(1, N + 1)
    lam1 = 2.0 - 2.0 * np.cos(k * np.pi / (N + 1))
    allv = (lam1[:, None, None] + lam1[None, :, None] + lam1[None, None, :]).ravel()
    return np.sort(allv)[:count]


def laplacian_3d_lowest_modes(N, count):
    """(i, j, k) mode indices of the `count` lowest eigenvalues of the N^3 Dirichlet Laplacian (ties in a fixed order)."""
    import numpy as np
    k = np.arange(1, N + 1)
    lam1 = 2.0 - 2.0 * np.cos(k * np.pi / (N + 1))
    allv = (lam1[:, None, None] + lam1[None, :, None] + lam1[None, None, :]).ravel()
    idx = np.argsort(allv, kind="stable")[:count]
    return np.stack(np.unravel_index(idx, (N, N, N)), axis=1) + 1


def analytic_subspace_angle(N, X, row0, nrows, allreduce):
    """sin of the largest angle between span(X) (this rank's rows [row0, row0 + nrows) of the computed eigenvectors, n x M) and the
    span of the M lowest analytic eigenvectors (sine products).  allreduce(array) sums a small array over the ranks in place."""
    import numpy as np
    M = X.shape[1]
    modes = laplacian_3d_lowest_modes(N, M)
    rows = np.arange(row0, row0 + nrows)
    x, y, z = rows // (N * N), (rows // N) % N, rows % N          # row = x N^2 + y N + z (Kronecker order of laplacian_3d)
    s = np.sqrt(2.0 / (N + 1)) * np.sin(np.outer(np.arange(1, N + 1), np.arange(1, N + 1)) * np.pi / (N + 1))   # s[mode-1, point]
    V = np.empty((nrows, M))
    for c, (i, j, k) in enumerate(modes):
        V[:, c] = s[i - 1, x] * s[j - 1, y] * s[k - 1, z]
    Xl = np.ascontiguousarray(X[row0:row0 + nrows])
    G = V.T @ Xl                                                   # V^T X, summed over the ranks
    allreduce(G)
    R = Xl - V @ G                                                 # component of X outside the analytic invariant subspace
    nr = np.array([np.sum(R * R, axis=0), np.sum(Xl * Xl, axis=0)])
    allreduce(nr)
    return float(np.sqrt((nr[0] / np.maximum(nr[1], 1e-300)).max()))


def workload(N, M0):
    import numpy as np
    A = laplacian_3d(N)
    ev = laplacian_3d_eigs(N, 80)
    Emin, Emax = 0.0, 0.5 * (ev[C3_M - 1] + ev[C3_M]) if N >= 12 else 0.5 * (ev[9] + ev[10])
    rng = np.random.default_rng(12345)
    Q0 = rng.standard_normal((N ** 3, M0))
    Q0 /= np.linalg.norm(Q0, axis=0)
    return A, ev, Emin, Emax, np.asfortranarray(Q0)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md, "clocks line")."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def cpu_sample(A, M0, steps, workers):
    """(seconds pass 1, seconds pass 2) of `steps` lock-step Lanczos steps on all M0 columns, `workers` processes."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import feast_port as fp
    return fp.time_mslanczos_sample(A, M0, steps, workers=workers)


def cpu_extrapolate(t1, t2, steps, lanczos_steps):
    return (t1 + t2) / steps * lanczos_steps


def reference_path_pair(eng, fc, N):
    """An honest end-to-end pair at a size the CPU finishes in seconds: the oracle's restatement of the reference's serial sparse path
    (oracle/feast_oracle.py:feast_scsrev = _feast_sparse_hermitian, sparse/feast_sparse.jl:246-499: sequential node loop, one sparse LU per
    node (SuperLU for UMFPACK), Rayleigh-Ritz, refinement) run TO COMPLETION beside the GPU engine on the same inputs.  Two CPU runs:
    the reference's own complex half-contour filter (it usually ends with info = 5 on this workload, the reference's behaviour) and the
    true filter rho = Re g the engine uses (same converged pairs)."""
    import numpy as np
    sys.path.insert(0, str(ROOT / "oracle"))
    import feast_oracle as fo
    A = laplacian_3d(N)
    ev = laplacian_3d_eigs(N, 80)
    want = C3_M if N >= 12 else 10
    Emin, Emax = 0.0, 0.5 * (ev[want - 1] + ev[want])
    M0 = 64 if N >= 12 else 26
    Q0 = np.random.default_rng(12345).standard_normal((N ** 3, M0))
    Q0 /= np.linalg.norm(Q0, axis=0)
    out = {"n": N ** 3, "M0": M0, "interval_holds": want, "cores": os.cpu_count() or 1}
    for name, filt in (("cpu_reference_filter", "reference"), ("cpu_true_filter", "true")):
        t0 = time.perf_counter()
        r = fo.feast_scsrev(A.tocsc(), Emin, Emax, M0, fo.feastinit(), Q0=Q0.astype(complex), solver="direct", filter=filt)
        out[name] = {"seconds": time.perf_counter() - t0, "info": int(r.info), "M": int(r.M), "loops": int(r.loop), "epsout": float(r.epsout)}
    t0 = time.perf_counter()
    rg = fc.feast_scsrev(A.tocsc(), Emin, Emax, M0, fc.feastinit(), Q0=Q0, solver_maxiter=3000, engine=eng)
    out["gpu"] = {"seconds": time.perf_counter() - t0, "info": int(rg.info), "M": int(rg.M), "loops": int(rg.loop), "epsout": float(rg.epsout),
                  "max_eig_err_vs_analytic": float(np.abs(np.sort(rg.lambda_) - ev[:rg.M]).max()) if rg.M else None}
    out["note"] = ("oracle = CPU restatement of FeastKit's :serial sparse path with a direct solver, run to completion; at 40^3 the same path "
                   "needs 767.5 s (:serial) / 128.2 s (:threads analog) on 8 cores against 73.5 ms on the GPU (tools/cpu_reference_path_40.py, BASELINE.md section 4), the GMRES path "
                   "returns info = 5; 100^3 is extrapolated by the 'port' entry")
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    A, ev, Emin, Emax, Q0 = workload(args.grid, args.m0)
    cores = os.cpu_count() or 1
    S = args.cpu_sample_steps
    times = []
    for i in range(args.warmup + args.steps):
        t1, t2 = cpu_sample(A, args.m0, S, cores)
        if i >= args.warmup:
            times.append(cpu_extrapolate(t1, t2, S, C3_LANCZOS_STEPS))
    sec = statistics.mean(times)
    val = C3_M / sec
    sample = (f"{S} lock-step Lanczos steps (pass 1 + pass 2 arithmetic) on all {args.m0} columns at n={args.grid ** 3}, "
              f"{cores} processes; extrapolated x{C3_LANCZOS_STEPS}/{S} to the {C3_LANCZOS_STEPS} steps of the converged solve "
              f"(Rayleigh-Ritz stages not included)")
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "config": config_dict(args),
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "note": "reference is pure Julia (no julia in the image): CPU port of the engine's algorithm (oracle/feast_port.py), "
                   "extrapolated from a bounded sample"}
    print(json.dumps(out))


def config_dict(args):
    return {"workload": f"configs[2]: sparse 3D 7-point Laplacian CSR n={args.grid ** 3} Float64, M0={args.m0}, 8 Gauss nodes, "
                        f"interval (0, mid(lambda_35, lambda_36)) -> M=35, fpm[3]=12",
            "grid": args.grid, "n": args.grid ** 3, "M0": args.m0, "nodes": 8, "inner_solver": "multi-shift two-pass Lanczos",
            "inner_rel": SOLVER_KW["inner_rel"], "fpm42": 0 if args.fp64 else 1, "parallelism": f"{args.shard if args.gpus > 1 else 'single GPU'} x{args.gpus}",
            "l2": "inputs larger than L2 (each block vector is 512 MB, L2 is 126 MB)"}


def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import feastcuda as fc

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if local == 0:
        sys.path.insert(0, str(ROOT))
        import __graft_entry__ as g
        g.build()   # no-op when the in-tree libfeastcuda.so is current
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (libfeastcuda has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    A, ev, Emin, Emax, Q0 = workload(args.grid, args.m0)
    eng = fc.default_engine(local)
    eng.set_sparse(fc.A, A, fc.SYM)
    eng.clear_b()
    eng.init_distributed()
    shard = args.shard if (world > 1 or os.environ.get("FEASTCUDA_FORCE_ROWS")) else "columns"
    eng.set_row_sharding(shard == "rows")
    fpm = fc.feastinit()
    fc.feastdefault_(fpm)
    Z, W = fc.feast_contour(Emin, Emax, fpm)
    # the headline runs what the reference-named API runs by default: fpm[42] = 1 ("single-precision solver", the reference's default,
    # core/feast_parameters.jl:316-319) -> FP32 Krylov vectors inside the FP64 refinement loop; --fp64 keeps FP64 vectors (fpm[42] = 0)
    head_mixed = False if args.fp64 else "fpm"
    opts = eng.make_opts(q0_real=True, x_real=True, shard=shard, mixed=head_mixed, **SOLVER_KW)
    # pinned host copies of the step's input (e2e leg)
    Q0_pinned = torch.from_numpy(Q0.T.copy()).pin_memory()   # (M0, n) C-order == (n, M0) column-major
    Q0_host = Q0_pinned.numpy().T
    if shard == "rows" and world > 1:
        # the first upload maps the peers' arenas (CUDA VMM handles over Unix sockets); a node without peer access fails on EVERY rank alike
        # (all-or-nothing inside the library) -> the run continues column-sharded instead of dying
        try:
            eng.upload_subspace(args.m0, Q0_host)
        except fc.FeastCudaError as exc:
            if rank == 0:
                print(f"bench.py: row sharding unavailable ({exc}); continuing with --shard columns", file=sys.stderr)
            shard = args.shard = "columns"
            eng.set_row_sharding(False)
            opts = eng.make_opts(q0_real=True, x_real=True, shard=shard, mixed=head_mixed, **SOLVER_KW)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def resident_step():
        eng.upload_subspace(args.m0, Q0_host)         # outside the timed region (inputs resident in HBM)
        barrier()
        s0 = eng.stats()["ms_dev_run"]
        M, info, eps, loop = eng.run_interval(Emin, Emax, args.m0, list(fpm), Z, W, opts)
        ms = eng.stats()["ms_dev_run"] - s0            # CUDA events on the library's stream around the whole solve
        return ms, M, info, eps, loop

    def e2e_step():
        barrier()
        t0 = time.perf_counter()
        r = eng.solve_interval(Emin, Emax, args.m0, list(fpm), Z, W, Q0=Q0_host, x_real=True, shard=shard, gather_rows=False, reuse_output=True, mixed=head_mixed, **SOLVER_KW)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3, r

    for _ in range(args.warmup):
        resident_step()
    eng.reset_stats()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    ms_list, last = [], None
    for _ in range(args.steps):
        ms, M, info, eps, loop = resident_step()
        ms_list.append(ms)
        last = (M, info, eps, loop)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    st = eng.stats()
    ms_step = sum(ms_list) / len(ms_list)
    # e2e: host buffers in, host buffers out, every step
    e2e_ms, r = [], None
    for i in range(1 + max(1, args.steps // 2)):
        ms, r = e2e_step()
        if i > 0:
            e2e_ms.append(ms)
    e2e_step_ms = sum(e2e_ms) / len(e2e_ms)
    lam, X, res = eng.fetch_results(args.m0, last[0], True)      # the FP64 solve's pairs, before the secondary leg overwrites them
    # secondary leg, reported beside the headline and never mixed into it: the same solve with the OTHER precision setting (FP64 Krylov
    # vectors when the headline follows fpm[42] = 1, FP32 vectors under --fp64)
    opts_fp64 = opts
    opts = eng.make_opts(q0_real=True, x_real=True, shard=shard, mixed=bool(args.fp64), **SOLVER_KW)
    mx_ms, mx_last, mx_error, st_mx, mx_step, mx_lam, mx_res = [], None, None, None, float("nan"), None, None
    try:      # the secondary leg must never cost the headline its JSON line
        if not args.no_mixed:
            resident_step()
        eng.reset_stats()
        for _ in range(1 if args.no_mixed else max(1, args.steps // 2 + 1)):
            ms, mM, minfo, meps, mloop = resident_step()
            mx_ms.append(ms)
            mx_last = (mM, minfo, meps, mloop)
        st_mx = eng.stats()
        mx_step = sum(mx_ms) / len(mx_ms)
        mx_lam, _, mx_res = eng.fetch_results(args.m0, mx_last[0], True)
    except Exception as exc:   # noqa: BLE001
        mx_error = f"{type(exc).__name__}: {exc}"
    opts = opts_fp64
    # ---- parity, on EVERY rank: M, eigenvalues vs the analytic spectrum, residuals, subspace angle vs the analytic eigenvectors
    M_loc, info_loc = last[0], last[1]
    row0, nrows, _ = eng.row_range()

    def allreduce_np(a):
        if world > 1 and shard == "rows":      # column-sharded ranks hold every row already
            tt = torch.from_numpy(a).cuda()
            dist.all_reduce(tt)
            a[...] = tt.cpu().numpy()
        return a

    angle = analytic_subspace_angle(args.grid, X, row0, nrows, allreduce_np) if M_loc == C3_M and args.grid >= 12 else float("nan")
    eig_err_loc = float(np.abs(np.sort(lam) - ev[:M_loc]).max()) if M_loc else float("inf")
    res_loc = float(res.max()) if M_loc else float("inf")
    ok_loc = float(info_loc == 0 and M_loc == (C3_M if args.grid >= 12 else 10) and eig_err_loc <= 1e-10 * max(1.0, float(ev[C3_M])) and res_loc < 1e-12
                   and (not np.isfinite(angle) or angle < 1e-8))
    worst = [eig_err_loc, res_loc, -ok_loc, float(M_loc), -float(M_loc)]
    if world > 1:
        wt = torch.tensor(worst, device="cuda", dtype=torch.float64)
        dist.all_reduce(wt, op=dist.ReduceOp.MAX)
        worst = wt.cpu().tolist()
    parity = {"all_ranks_ok": bool(-float(worst[2]) == 1.0), "same_M_on_all_ranks": bool(float(worst[3]) == -float(worst[4])),
              "max_eig_err_vs_analytic_over_ranks": float(worst[0]), "max_residual_over_ranks": float(worst[1]),
              "subspace_angle_vs_analytic": angle, "ranks": world}
    if world > 1:
        t = torch.tensor([ms_step, e2e_step_ms, mx_step], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, e2e_step_ms, mx_step = float(t[0]), float(t[1]), float(t[2])
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    M, info, eps, loop = last
    eig_err = float(np.abs(np.sort(lam) - ev[:M]).max()) if M else None
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
    names = fc._lib.KERN_NAMES

    def kernel_table(stx):
        tab = {}
        for i, nm in enumerate(names):
            if stx["n_kern"][i]:
                avg = stx["ms_kern"][i] / stx["n_kern"][i]
                tab[nm] = {"avg_ms": avg, "alg_bytes": stx["bytes_kern"][i], "gbs": stx["bytes_kern"][i] / avg / 1e6,
                           "frac_of_peak": stx["bytes_kern"][i] / avg / 1e6 / peak, "sampled": stx["n_kern"][i]}
        return tab

    kern = kernel_table(st)
    # total time share of each kernel kind over the timed steps: launches x average duration
    launches = {"lz_p1": st["lz_steps_p1"], "lz_upd": st["lz_steps_p1"], "lz_p2": st["lz_steps_p2"]}
    share = {k: launches[k] * kern[k]["avg_ms"] for k in launches if k in kern}
    dom = max(share, key=share.get) if share else None
    traffic = None
    prof = ROOT / "profiles" / "ncu_summary.json"
    fp32_run = st.get("lz_steps_fp32", 0) * 2 > st["lz_steps_p1"]
    if prof.exists() and dom and world == 1:          # the captures are single-GPU launches over all 64 columns
        summ = json.loads(prof.read_text())
        # pass 2 alternates two launch kinds (accumulating / skipping): its per-launch traffic is their average;
        # the FP32-vector kernels have their own captures (r2_lz32_*)
        fam = "lz32" if fp32_run else "lz"
        key = {"lz_p2": fam + "_p2_paired_average", "lz_p1": "r2_lz32_p1" if fp32_run else "lz_p1",
               "lz_upd": "r2_lz32_upd" if fp32_run else "lz_upd"}.get(dom, dom)
        traffic = summ.get(key, {}).get("dram_bytes_per_launch")
    roofline = None
    if dom:
        kfam = "k_lz32_spmm" if fp32_run else "k_lz_spmm"
        roofline = {"bound": "hbm", "kernel": {"lz_p1": kfam + "<LZ_P1>", "lz_p2": kfam + "<LZ_P2_PAIR|LZ_P2_SKIP> (pass 2, average launch)",
                                               "lz_upd": kfam.replace("spmm", "update")}[dom],
                    "achieved": kern[dom]["gbs"], "peak": peak, "unit": "GB/s", "frac": kern[dom]["gbs"] / peak, "traffic": traffic,
                    "peak_source": peak_src, "alg_bytes_per_launch": kern[dom]["alg_bytes"], "avg_launch_ms": kern[dom]["avg_ms"],
                    "share_of_step": share[dom] / (ms_step * args.steps), "all_kernels": kern}
    cores = os.cpu_count() or 1
    cpu = None
    if world == 1 and not args.no_cpu:
        S = args.cpu_sample_steps
        t1, t2 = cpu_sample(A, args.m0, S, cores)
        lz = st["lz_steps_p1"] / args.steps if st["lz_steps_p1"] else C3_LANCZOS_STEPS
        sec = cpu_extrapolate(t1, t2, S, lz)
        cpu = {"value": M / sec if sec > 0 else None, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{S} lock-step Lanczos steps (both passes) on {args.m0} columns at n={args.grid ** 3} with {cores} processes "
                         f"({t1 + t2:.1f} s), extrapolated to the {lz:.0f} steps this solve took; oracle/feast_port.py"}
    if cpu is not None:
        cpu["reference_path_pair"] = reference_path_pair(eng, fc, args.cpu_pair_grid)
    out = {"metric": METRIC, "value": M / (ms_step / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f64" if args.fp64 else "f64 outer loop / f32 Krylov vectors (fpm[42] = 1, the reference's default)", "data": "synthetic",
           "config": config_dict(args),
           "result": {"M": M, "info": info, "epsout": eps, "loops": loop, "max_residual": float(res.max()) if M else None,
                      "max_eig_err_vs_analytic": eig_err, "subspace_angle_vs_analytic": parity["subspace_angle_vs_analytic"],
                      "lanczos_steps_per_solve": st["lz_steps_p1"] / args.steps, "fp32_steps_per_solve": st.get("lz_steps_fp32", 0) / args.steps,
                      "parity": parity},
           "e2e": {"value": r.M / (e2e_step_ms / 1e3), "unit": UNIT, "ms_per_step": e2e_step_ms,
                   "h2d_bytes_per_step": int(Q0.nbytes) if shard == "rows" or world == 1 else int(Q0.nbytes) * world,
                   "d2h_bytes_per_step": (int(r.q.nbytes) if shard == "rows" or world == 1 else int(r.q.nbytes) * world) + int(r.lambda_.nbytes + r.res.nbytes) * world,
                   "operator_h2d_bytes_once": int(A.nnz * 12 + 4 * (A.shape[0] + 1)),     # CSR upload at set_csr (int32 indices), outside the timed region
                   "note": "whole-job bytes: row-sharded ranks copy their own rows of Q0 / X only, column-sharded ranks the full blocks"},
           "gpu_launches": int(st["kernel_launches"]), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
           "mixed_precision": None if args.no_mixed else ({"error": mx_error} if mx_error else {
               "what": ("same solve with FP32 Lanczos vectors and matrix entries (fpm[42] = 1), FP64 scalars, accumulator, Rayleigh-Ritz and residuals"
                        if args.fp64 else "same solve with FP64 Krylov vectors throughout (fpm[42] = 0 / mixed=False)") + "; not the headline",
               "ms_per_step": mx_step, "value": mx_last[0] / (mx_step / 1e3), "unit": UNIT,
               "result": {"M": mx_last[0], "info": mx_last[1], "epsout": mx_last[2], "loops": mx_last[3],
                          "max_residual": float(mx_res.max()) if mx_last[0] else None,
                          "max_eig_err_vs_analytic": float(np.abs(np.sort(mx_lam) - ev[:mx_last[0]]).max()) if mx_last[0] else None,
                          "lanczos_steps_per_solve": st_mx["lz_steps_p1"] / len(mx_ms), "fp32_steps_per_solve": st_mx["lz_steps_fp32"] / len(mx_ms)},
               "kernels": kernel_table(st_mx)})}
    print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# =====================================================================================================================
# the other BASELINE configs at full size on ONE GPU (not the headline): a line of the same shape, e2e through the reference-named API
# =====================================================================================================================
def fem_pair(nx, ny, nz, seed=7):
    """configs[3]: trilinear-FEM stiffness/mass Kronecker pair under a unitary diagonal gauge (complex Hermitian)."""
    import numpy as np
    import scipy.sparse as sp

    def k1(n, h):
        return sp.diags([-np.ones(n - 1), 2 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1]) / h

    def m1(n, h):
        return h * sp.diags([np.ones(n - 1), 4 * np.ones(n), np.ones(n - 1)], [-1, 0, 1]) / 6
    hs = [1.0 / (n + 1) for n in (nx, ny, nz)]
    Ks = [k1(n, h) for n, h in zip((nx, ny, nz), hs)]
    Ms = [m1(n, h) for n, h in zip((nx, ny, nz), hs)]
    K = sp.kron(sp.kron(Ks[0], Ms[1]), Ms[2]) + sp.kron(sp.kron(Ms[0], Ks[1]), Ms[2]) + sp.kron(sp.kron(Ms[0], Ms[1]), Ks[2])
    Mass = sp.kron(sp.kron(Ms[0], Ms[1]), Ms[2])
    n = nx * ny * nz
    D = sp.diags(np.exp(1j * np.random.default_rng(seed).uniform(0, 2 * np.pi, n)))
    A = (D @ K @ D.conj()).tocsc()
    B = (D @ Mass @ D.conj()).tocsc()
    A = ((A + A.conj().T) * 0.5).tocsc()
    B = ((B + B.conj().T) * 0.5).tocsc()
    lam = []
    for n_, h in zip((nx, ny, nz), hs):
        th = np.arange(1, n_ + 1) * np.pi / (n_ + 1)
        lam.append(((2 - 2 * np.cos(th)) / h) / (h * (4 + 2 * np.cos(th)) / 6))
    w = np.sort((lam[0][:, None, None] + lam[1][None, :, None] + lam[2][None, None, :]).ravel())
    return A, B, w


def toeplitz_pencil(dims, eps=0.05):
    """configs[4]: Kronecker sums of non-symmetric complex Toeplitz tridiagonals, B = I + eps * (same b/c ratio)."""
    import numpy as np
    import scipy.sparse as sp
    coef = [(0.4 + 0.1j, 1.0 + 0.05j, 0.9 - 0.05j), (0.3 - 0.1j, 0.8 + 0.1j, 0.75 + 0.05j), (0.5 + 0.2j, 0.6 - 0.05j, 0.65 + 0.02j)]
    toe = lambda n, a, b, c: sp.diags([b * np.ones(n - 1), a * np.ones(n), c * np.ones(n - 1)], [-1, 0, 1])
    I = [sp.identity(n) for n in dims]
    T = [toe(n, *abc) for n, abc in zip(dims, coef)]
    S = [toe(n, 0.0, abc[1], abc[2]) for n, abc in zip(dims, coef)]
    ksum = lambda X: sp.kron(sp.kron(X[0], I[1]), I[2]) + sp.kron(sp.kron(I[0], X[1]), I[2]) + sp.kron(sp.kron(I[0], I[1]), X[2])
    A = ksum(T).tocsc()
    B = (sp.identity(A.shape[0]) + eps * ksum(S)).tocsc()
    la, ls = [], []
    for n, (a, b, c) in zip(dims, coef):
        th = np.arange(1, n + 1) * np.pi / (n + 1)
        la.append(a + 2 * np.sqrt(b * c) * np.cos(th))
        ls.append(2 * np.sqrt(b * c) * np.cos(th))
    lamA = (la[0][:, None, None] + la[1][None, :, None] + la[2][None, None, :]).ravel()
    lamS = (ls[0][:, None, None] + ls[1][None, :, None] + ls[2][None, None, :]).ravel()
    return A, B, lamA / (1 + eps * lamS)


def band_workload(N, kb=7, c=0.3, want=40, centre=2.5):
    """The banded line (not a BASELINE config): A = T_N (x) I_kb + I_N (x) c T_kb with the kb-index fastest -- a real symmetric
    band matrix of order kb N and half-bandwidth kb whose eigenvalues t_i + mu_a are analytic.  The interval sits in the middle of the
    spectrum, its ends in the widest gaps near the targets."""
    import numpy as np
    import scipy.sparse as sp
    TN = sp.diags([-np.ones(N - 1), 2 * np.ones(N), -np.ones(N - 1)], [-1, 0, 1])
    D = c * sp.diags([-np.ones(kb - 1), 2 * np.ones(kb), -np.ones(kb - 1)], [-1, 0, 1])
    A = (sp.kron(TN, sp.identity(kb)) + sp.kron(sp.identity(N), D)).tocsc()
    t = 2 - 2 * np.cos(np.arange(1, N + 1) * np.pi / (N + 1))
    mu = c * (2 - 2 * np.cos(np.arange(1, kb + 1) * np.pi / (kb + 1)))
    lam = np.sort((t[:, None] + mu[None, :]).ravel())
    i0 = int(np.searchsorted(lam, centre))
    widest = lambda i: max(range(i - 4, i + 5), key=lambda q: lam[q] - lam[q - 1])
    lo, hi = widest(i0), widest(i0 + want)
    AB = np.zeros((kb + 1, kb * N))                      # LAPACK upper band storage, diagonal in row kb (banded/feast_banded.jl:205-214)
    for d in range(kb + 1):
        AB[kb - d, d:] = A.diagonal(d)
    return A, AB, lam[lo:hi], 0.5 * (lam[lo - 1] + lam[lo]), 0.5 * (lam[hi - 1] + lam[hi])


def fp64_gemm_peak(n=8192, reps=3):
    """Denominator only: the library ZGEMM / DGEMM rate of this box (torch.matmul -> cuBLAS), TFLOP/s (real flops)."""
    import torch
    out = {}
    for name, dt, fl in (("zgemm", torch.complex128, 8.0), ("dgemm", torch.float64, 2.0)):
        a = torch.randn(n, n, dtype=dt, device="cuda")
        b = torch.randn(n, n, dtype=dt, device="cuda")
        torch.matmul(a, b)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = float("inf")
        for _ in range(reps):
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out[name] = fl * n ** 3 / (best * 1e-3) / 1e12
        del a, b
    return out


def run_other_config(args):
    import numpy as np
    import torch
    import feastcuda as fc
    sys.path.insert(0, str(ROOT))
    import __graft_entry__ as g
    g.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (libfeastcuda has no CPU fallback)")
    torch.cuda.set_device(0)
    rng = np.random.default_rng(12345)
    cfg = args.config
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    if cfg == 1:
        n, M0 = args.n or 8192, 128
        v = np.random.default_rng(42).standard_normal(n)
        v /= np.linalg.norm(v)
        d = np.linspace(0.0, 100.0, n)
        Dv = d * v
        A = np.diag(d) - 2 * np.outer(v, Dv) - 2 * np.outer(Dv, v) + 4 * (v @ Dv) * np.outer(v, v)
        A = 0.5 * (A + A.T)
        Emin, Emax = 50.0, 50.0 + 80.5 * (100.0 / (n - 1))
        exact = d[(d >= Emin) & (d <= Emax)]
        Q0 = rng.standard_normal((n, M0))
        Q0 /= np.linalg.norm(Q0, axis=0)
        solve = lambda: fc.dfeast_syev(A, Emin, Emax, M0, fc.feastinit(), Q0=Q0)
        check = lambda r: float(np.abs(np.sort(r.lambda_) - exact).max()) if r.M == len(exact) else None
        workload = f"configs[1]: dense real symmetric dfeast_syev! n={n} Float64 (Householder-similar to diag(linspace(0,100,n))), M0=128, 8 Gauss nodes"
        # flops of one solve: 8 complex LU factorisations (8/3 n^3 each) + per loop 8 x (two triangular block solves 8 n^2 m) + A*Q
        flops = lambda r: 8 * (8.0 / 3.0) * n ** 3 + (r.loop + 1) * (8 * 8.0 * n * n * M0 + 8.0 * n * n * M0)
        h2d = A.nbytes + Q0.nbytes
    elif cfg == 3:
        dims = (args.n3 or (100, 100, 50))
        A, B, w = fem_pair(*dims)
        n, M0, want = A.shape[0], 96, 60
        while w[want] - w[want - 1] < 1e-8 * w[want]:
            want += 1
        Emin, Emax = 0.0, 0.5 * (w[want - 1] + w[want])
        Q0 = rng.standard_normal((n, M0)) + 0j
        Q0 /= np.linalg.norm(Q0, axis=0)

        def solve():
            fpm = fc.feastinit()
            fpm[1] = 16
            return fc.zfeast_hcsrgv(A, B, Emin, Emax, M0, fpm, Q0=Q0, solver_maxiter=6000)
        check = lambda r: float(np.abs(np.sort(r.lambda_) - w[:want]).max() / w[want]) if r.M == want else None
        exact = w[:want]
        workload = f"configs[3]: zfeast_hcsrgv! trilinear-FEM stiffness/mass pair {dims[0]}x{dims[1]}x{dims[2]} (n={n}), complex Hermitian, M0=96, 16 nodes, lowest {want} pairs"
        flops = None
        h2d = Q0.nbytes + A.data.nbytes + A.indices.nbytes + B.data.nbytes + B.indices.nbytes
    elif cfg == 5:
        kb, Nb = 7, (args.n or 999999) // 7
        Asp, AB, exact, Emin, Emax = band_workload(Nb, kb)
        n, M0 = kb * Nb, 64
        Q0 = rng.standard_normal((n, M0))
        Q0 /= np.linalg.norm(Q0, axis=0)
        solve = lambda: fc.dfeast_sbev(AB, kb, Emin, Emax, M0, fc.feastinit(), Q0=Q0)
        check = lambda r: float(np.abs(np.sort(r.lambda_) - exact).max()) if r.M == len(exact) else None
        workload = (f"banded path (SURVEY 8a a14/a15; not a BASELINE config): dfeast_sbev! real symmetric band n={n}, half-bandwidth {kb} "
                    f"(T_N (x) I_7 + I_N (x) 0.3 T_7, analytic spectrum), M0=64, 8 Gauss nodes, {len(exact)} eigenvalues around 2.5 "
                    f"(interval width {Emax - Emin:.2e}); one band LU per node (all nodes in one launch), cached across the refinement loops")
        flops = None
        h2d = AB.nbytes + Q0.nbytes
    else:
        dims = (args.n3 or (50, 50, 100))
        A, B, lam = toeplitz_pencil(dims)
        n, M0 = A.shape[0], 64
        order = np.argsort(lam.real)
        Emid = complex(lam[order[0]].real, lam[order[:40]].imag.mean())
        dist = np.abs(lam - Emid)
        rad = 0.5 * (np.sort(dist)[34] + np.sort(dist)[35])
        exact = lam[dist <= rad]
        Q0 = rng.standard_normal((n, M0)) + 0j
        Q0 /= np.linalg.norm(Q0, axis=0)

        def solve():
            fpm = fc.feastinit()
            fpm[7], fpm[2], fpm[3] = 24, 10, 30
            return fc.pzifeast_gcsrgv(A, B, Emid, rad, M0, fpm, Q0=Q0, solver_maxiter=4000)
        check = lambda r: float(max(min(abs(gv - x) for x in exact) for gv in r.lambda_)) if r.M == len(exact) else None
        workload = (f"configs[4]: pzifeast_gcsrgv! general complex Toeplitz-Kronecker pencil {dims[0]}x{dims[1]}x{dims[2]} (n={n}), M0=64, 24 nodes, "
                    f"disc at the left edge of the spectrum holding 35 eigenvalues (radius {rad:.4f}; fpm[3]=10)")
        flops = None
        h2d = Q0.nbytes + A.data.nbytes + A.indices.nbytes + B.data.nbytes + B.indices.nbytes
    times, r = [], None
    sampler = ClockSampler(0)
    for i in range(args.warmup + args.steps):
        if i == args.warmup:
            sampler.start()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = solve()
        torch.cuda.synchronize()
        if i >= args.warmup:
            times.append((time.perf_counter() - t0) * 1e3)
    clocks = sampler.stop()
    ms = sum(times) / len(times)
    st = r.stats
    names = fc._lib.KERN_NAMES
    kern = {}
    for i, nm in enumerate(names):
        if st["n_kern"][i]:
            avg = st["ms_kern"][i] / st["n_kern"][i]
            kern[nm] = {"avg_ms": avg, "alg_bytes": st["bytes_kern"][i], "gbs": st["bytes_kern"][i] / avg / 1e6,
                        "frac_of_peak": st["bytes_kern"][i] / avg / 1e6 / hbm, "sampled": st["n_kern"][i]}
    if cfg == 1:
        pk = fp64_gemm_peak()
        ach = flops(r) / (st["ms_total"] * 1e-3) / 1e12
        roofline = {"bound": "fp64-tensor", "kernel": "k_zgemm_dmma_async + LU panels (whole solve)", "achieved": ach, "peak": pk["zgemm"], "unit": "TFLOP/s",
                    "frac": ach / pk["zgemm"], "traffic": None, "peak_source": "measured in this run: torch.matmul complex128 8192^3 (cuBLAS ZGEMM), denominator only",
                    "dgemm_tflops": pk["dgemm"], "flops_counted": flops(r)}
    else:
        dom = max(kern, key=lambda k_: kern[k_]["avg_ms"] * kern[k_]["sampled"]) if kern else None
        roofline = None if dom is None else {"bound": "hbm", "kernel": {"lz_cheb": "k_lz_spmm<LZ_CHEB> (Chebyshev / Jacobi step of the inner solve with B)",
                                                                          "lz_p1": "k_lz_spmm<LZ_P1> (A times the Lanczos block)",
                                                                          "band_lu": "k_band_lu_smem (band LU, one warp per node, window in shared memory; n dependent steps: instruction-issue-bound)",
                                                                          "band_solve": "k_band_solve_lanes (band substitutions of all nodes x columns in one launch, window spread over lanes; n dependent steps: instruction-issue-bound)"}.get(dom, dom),
                                             "achieved": kern[dom]["gbs"], "peak": hbm, "unit": "GB/s", "frac": kern[dom]["gbs"] / hbm, "traffic": None,
                                             "alg_bytes_per_launch": kern[dom]["alg_bytes"], "avg_launch_ms": kern[dom]["avg_ms"], "all_kernels": kern}
    cpu = None
    if cfg == 5 and not args.no_cpu:
        # the reference's :serial banded path restated (sequential node loop, one LAPACK band LU per node -- zgbtrf/zgbtrs, cached; Rayleigh-Ritz;
        # refinement; banded/feast_banded.jl:561-823) run TO COMPLETION on the same inputs
        sys.path.insert(0, str(ROOT / "oracle"))
        import feast_oracle as fo
        t0 = time.perf_counter()
        ro = fo.feast_hrr(Asp.astype(complex), None, Emin, Emax, M0, fo.feastinit(), Q0=Q0.astype(complex), filter="true", band_k=kb)
        sec = time.perf_counter() - t0
        cpu = {"value": ro.M / sec, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
               "sample": f"the whole solve at full size, run to completion: oracle/feast_oracle.py feast_hrr (LAPACK band LU per node, true filter), {sec:.1f} s, "
                         f"info={ro.info}, M={ro.M}, loops={ro.loop}, epsout={ro.epsout:.2e}",
               "max_eig_diff_gpu_vs_cpu": float(np.abs(np.sort(ro.lambda_.real) - np.sort(r.lambda_)).max()) if ro.M == r.M else None}
    name = "banded path, n=10^6 k=7" if cfg == 5 else f"BASELINE configs[{cfg}] at full size"
    out = {"metric": f"FEAST solve eigenpairs/s, {name} (wall-time in ms_per_step)", "value": r.M / (ms / 1e3), "unit": UNIT,
           "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "config": {"workload": workload, "n": int(n), "M0": M0},
           "result": {"M": r.M, "expected_M": int(len(exact)), "info": r.info, "epsout": r.epsout, "loops": r.loop, "max_residual": float(r.res.max()) if r.M else None,
                      "max_eig_err_vs_analytic": check(r), "lanczos_steps": st["lz_steps_p1"], "inner_solve_degree": st.get("cheb_degree", 0)},
           "e2e": {"value": r.M / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(r.q.nbytes + r.lambda_.nbytes + r.res.nbytes),
                   "note": "timed through the reference-named API with host matrices and host Q0 (operator upload included); device-resident time: result of feastcuda_stats.ms_total below"},
           "device_ms_total": st["ms_total"], "gpu_launches": int(st["kernel_launches"]), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", type=int, default=100, help="grid points per dimension (n = grid^3)")
    ap.add_argument("--m0", type=int, default=64)
    ap.add_argument("--cpu-sample-steps", type=int, default=8)
    ap.add_argument("--cpu-pair-grid", type=int, default=16, help="grid of the end-to-end CPU-oracle / GPU pair (n = grid^3)")
    ap.add_argument("--shard", default="rows", choices=["rows", "columns"], help="multi-GPU partition of the Lanczos filter (N > 1)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-mixed", action="store_true", help="skip the secondary leg with the other precision setting (profiling runs)")
    ap.add_argument("--fp64", action="store_true", help="headline with FP64 Krylov vectors (fpm[42] = 0) instead of the reference's default fpm[42] = 1")
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5], help="index into BASELINE.json's configs (2 = the headline workload); 5 = the banded path (n = 10^6, k = 7)")
    ap.add_argument("--n", type=int, default=0, help="--config 1: matrix order (default 8192); --config 5: matrix order (default 999999)")
    ap.add_argument("--n3", type=int, nargs=3, default=None, help="--config 3|4: grid (default: the full size)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.config != 2:
        if "--steps" not in sys.argv:
            args.steps = 1
        if "--warmup" not in sys.argv:
            args.warmup = 1 if args.config in (1, 5) else 0
        run_other_config(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
