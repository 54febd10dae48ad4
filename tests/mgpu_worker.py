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
    for name, kw in (("mslanczos_columns", dict(solver="mslanczos", solver_maxiter=2000)),
                     ("bicgstab_nodes", dict(solver="bicgstab", solver_maxiter=400, inner_rel=1e-3, ritz_guess=True, shard="nodes")),
                     ("bicgstab_balanced", dict(solver="bicgstab", solver_maxiter=400, inner_rel=1e-3, ritz_guess=True, shard="balanced"))):
        r = fc.pdfeast_scsrev(A, Emin, Emax, M0, fc.feastinit(), Q0=Q0, **kw)
        ok = (r.info == 0 and r.M == 10 and float(np.abs(np.sort(r.lambda_) - ev[:10]).max()) < 1e-10 and float(r.res.max()) < 1e-12)
        t = torch.tensor([1.0 if ok else 0.0, float(r.loop)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        out[name] = {"ok_all_ranks": bool(t[0].item() == 1.0), "loop": r.loop, "M": r.M, "info": r.info, "epsout": r.epsout,
                     "allreduce_bytes": r.stats["allreduce_bytes"], "lz_steps": r.stats["lz_steps_p1"], "world": world}
    if rank == 0:
        Path(sys.argv[1]).write_text(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
