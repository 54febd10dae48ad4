"""Worker for the multi-GPU parity test (launched by torchrun, one rank per GPU): solves a reduced config 3 with the
column-sharded Lanczos filter and with node-sharded BiCGStab, checks every rank against analytic eigenvalues, rank 0 writes a JSON."""
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "feastkit.jl_b200"))
sys.path.insert(0, str(ROOT / "oracle"))


def main():
    import torch
    import torch.distributed as dist
    import feast_oracle as fo
    import feastcuda as fc
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    N, M0 = 24, 26
    A = fo.laplacian_3d(N).astype(float).tocsc()
    ev = fo.laplacian_3d_eigs(N)
    Emin, Emax = 0.0, 0.5 * (ev[9] + ev[10])
    Q0 = fo.seeded_subspace(N ** 3, M0, complex_storage=False)
    out = {}
    for name, kw in (("mslanczos_rows", dict(solver="mslanczos", solver_maxiter=2000, shard="rows")),
                     ("mslanczos_columns", dict(solver="mslanczos", solver_maxiter=2000)),
                     ("mslanczos_mixed_columns", dict(solver="mslanczos", solver_maxiter=2000, mixed=True)),
                     ("bicgstab_nodes", dict(solver="bicgstab", solver_maxiter=400, inner_rel=1e-3, ritz_guess=True, shard="nodes")),
                     ("bicgstab_balanced", dict(solver="bicgstab", solver_maxiter=400, inner_rel=1e-3, ritz_guess=True, shard="balanced"))):
        r = fc.pdfeast_scsrev(A, Emin, Emax, M0, fc.feastinit(), Q0=Q0, **kw)
        ok = (r.info == 0 and r.M == 10 and float(np.abs(np.sort(r.lambda_) - ev[:10]).max()) < 1e-10 and float(r.res.max()) < 1e-12)
        if ok:      # the returned vectors are the full eigenvectors on every rank (row-sharded runs gather them)
            R = A @ r.q - r.q * r.lambda_
            ok = float(np.linalg.norm(R, axis=0).max()) < 1e-11 and float(np.abs(np.linalg.norm(r.q, axis=0) - 1).max()) < 1e-10
        t = torch.tensor([1.0 if ok else 0.0, float(r.loop)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        out[name] = {"ok_all_ranks": bool(t[0].item() == 1.0), "loop": r.loop, "M": r.M, "info": r.info, "epsout": r.epsout,
                     "allreduce_bytes": r.stats["allreduce_bytes"], "lz_steps": r.stats["lz_steps_p1"], "world": world}
    # matrix-free operator (compiled CUDA stencil callback), columns sharded like the CSR Lanczos path
    import ctypes as C

    class LaplacianGrid(C.Structure):
        _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int)]

    ex = C.CDLL(str(ROOT / "feastkit.jl_b200" / "lib" / "libfeastcuda_examples.so"))
    grid = LaplacianGrid(N, N, N)
    r = fc.feast_matvec((ex.feastcuda_example_laplacian3d, grid), None, N ** 3, (Emin, Emax), M0=M0, fpm=fc.feastinit(), Q0=Q0,
                        solver_maxiter=2000)
    ok = (r.info == 0 and r.M == 10 and float(np.abs(np.sort(r.lambda_) - ev[:10]).max()) < 1e-10 and float(r.res.max()) < 1e-12)
    t = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    out["matfree_columns"] = {"ok_all_ranks": bool(t[0].item() == 1.0), "loop": r.loop, "M": r.M, "info": r.info, "epsout": r.epsout,
                              "allreduce_bytes": r.stats["allreduce_bytes"], "lz_steps": r.stats["lz_steps_p1"], "world": world}
    # dense (LU per node, nodes sharded by the reference's block rule) and general (full contour) solves
    rng = np.random.default_rng(3)
    n = 160
    Qm, _ = np.linalg.qr(rng.standard_normal((n, n)))
    d = np.sort(rng.uniform(0.0, 10.0, n))
    Ad = (Qm * d) @ Qm.T
    Ad = 0.5 * (Ad + Ad.T)
    r = fc.pdfeast_syev(Ad, 0.5 * (d[29] + d[30]), 0.5 * (d[37] + d[38]), 20, fc.feastinit(), Q0=fo.seeded_subspace(n, 20, complex_storage=False))
    ok = r.info == 0 and r.M == 8 and float(np.abs(np.sort(r.lambda_) - d[30:38]).max()) < 1e-10 * d[-1] and float(r.res.max()) < 1e-12
    t = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    out["dense_nodes"] = {"ok_all_ranks": bool(t[0].item() == 1.0), "loop": r.loop, "M": r.M, "info": r.info, "epsout": r.epsout,
                          "allreduce_bytes": r.stats["allreduce_bytes"], "world": world}
    nt = 120
    a_, b_, c_ = 0.3 + 0.2j, 1.0 + 0.1j, 0.95 - 0.05j
    At = np.diag(b_ * np.ones(nt - 1), -1) + np.diag(a_ * np.ones(nt)) + np.diag(c_ * np.ones(nt - 1), 1)
    lam = a_ + 2 * np.sqrt(b_ * c_) * np.cos(np.arange(1, nt + 1) * np.pi / (nt + 1))
    Emid, rad = 0.3 + 0.2j, 0.3
    inside = [l for l in lam if abs(l - Emid) <= rad]
    r = fc.pzfeast_geev(At, Emid, rad, 24, fc.feastinit(), Q0=fo.seeded_subspace(nt, 24))
    ok = r.info == 0 and r.M == len(inside) and all(min(abs(g - x) for x in inside) < 1e-9 for g in r.lambda_) and float(r.res.max()) < 1e-12
    t = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    out["general_nodes"] = {"ok_all_ranks": bool(t[0].item() == 1.0), "loop": r.loop, "M": r.M, "info": r.info, "epsout": r.epsout,
                            "allreduce_bytes": r.stats["allreduce_bytes"], "world": world}
    if rank == 0:
        Path(sys.argv[1]).write_text(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
