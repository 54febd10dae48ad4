"""CPU, world_size 2 over gloo: the multi-GPU plan of the engine (DESIGN.md "Multi-GPU") exercised with the NumPy ports.

Lanczos path: every rank filters a contiguous slice of column PAIRS for all nodes; BiCGStab path: every rank owns a
contiguous block of quadrature nodes (parallel/feast_mpi.jl:36-43).  Either way the n x M0 accumulator is summed with ONE
all-reduce per refinement loop (MPI.Allreduce of Q_proj, parallel/feast_mpi.jl:119,341,858) and the small Rayleigh-Ritz
stage is replicated.  The sharded runs must give the single-rank result."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    """A TCP port nobody listens on right now (fixed pid-derived ports collided with other processes once in a while)."""
    import socket
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def pair_slice(active, nranks, rank):
    """Column slice owned by `rank`: contiguous column pairs (run_interval in csrc/feastcuda.cu)."""
    npairs = (active + 1) // 2
    pb, pr = divmod(npairs, nranks)
    p0 = rank * pb + min(rank, pr)
    pn = pb + (1 if rank < pr else 0)
    c0 = 2 * p0
    return c0, max(0, min(active, 2 * (p0 + pn)) - c0)


def _worker(rank, world, port, out):
    sys.path.insert(0, str(ROOT / "oracle"))
    import torch
    import torch.distributed as dist
    import feast_oracle as fo
    import feast_port as fp
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    N, M0 = 8, 13
    A = fo.laplacian_3d(N).astype(float).tocsr()
    ev = fo.laplacian_3d_eigs(N)
    Emin, Emax = 0.0, 0.5 * (ev[6] + ev[7])
    Q0 = fo.seeded_subspace(N ** 3, M0, complex_storage=False)

    def allreduce(acc):
        t = torch.from_numpy(np.ascontiguousarray(acc.view(np.float64) if np.iscomplexobj(acc) else acc))
        dist.all_reduce(t)
        return acc

    r = fp.feast_hrr_mslanczos(A, Emin, Emax, M0, fo.feastinit(), Q0, inner_rel=1e-3, inner_maxiter=800,
                               col_slices=lambda active: pair_slice(active, world, rank), allreduce=allreduce)
    # fpm[42] mixed precision: FP32 Lanczos vectors on every rank's slice, same FP64 all-reduce and Rayleigh-Ritz stage
    rm = fp.feast_hrr_mslanczos(A, Emin, Emax, M0, fo.feastinit(), Q0, inner_rel=1e-3, inner_maxiter=800, adaptive=True, mixed=True,
                                col_slices=lambda active: pair_slice(active, world, rank), allreduce=allreduce)
    ne = 8
    s, c = fo.node_partition(ne, world, rank)
    rb = fp.feast_hrr_bicgstab(A, None, Emin, Emax, M0, fo.feastinit(), Q0, inner_rel=1e-3, inner_maxiter=400,
                               node_items=lambda loop, active: [(e, 0, active) for e in range(s, s + c)], allreduce=allreduce)
    if rank == 0:
        np.savez(out, lam=np.sort(r.lambda_), M=r.M, info=r.info, loop=r.loop, q=r.q, lam_b=np.sort(rb.lambda_), M_b=rb.M,
                 info_b=rb.info, lam_m=np.sort(rm.lambda_), M_m=rm.M, info_m=rm.info, res_m=rm.res.max(), fp32_sweeps=rm.stats["fp32_sweeps"])
    dist.barrier()
    dist.destroy_process_group()


def test_pair_slices_tile_the_columns():
    for active in (1, 2, 7, 13, 35, 64, 128):
        for P in (1, 2, 3, 4, 8):
            cols = []
            for rank in range(P):
                c0, nc = pair_slice(active, P, rank)
                assert c0 % 2 == 0 and nc >= 0
                cols += list(range(c0, c0 + nc))
            assert cols == list(range(active))


@pytest.mark.timeout(600)
def test_world_size_2_sharded_solves_equal_single_rank(tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, str(ROOT / "oracle"))
    import feast_oracle as fo
    import feast_port as fp
    out = str(tmp_path / "r0.npz")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    N, M0 = 8, 13
    A = fo.laplacian_3d(N).astype(float).tocsr()
    ev = fo.laplacian_3d_eigs(N)
    Emin, Emax = 0.0, 0.5 * (ev[6] + ev[7])
    Q0 = fo.seeded_subspace(N ** 3, M0, complex_storage=False)
    r1 = fp.feast_hrr_mslanczos(A, Emin, Emax, M0, fo.feastinit(), Q0, inner_rel=1e-3, inner_maxiter=800)
    assert int(got["info"]) == r1.info == 0 and int(got["M"]) == r1.M == 7
    assert int(got["loop"]) == r1.loop
    assert np.abs(got["lam"] - np.sort(r1.lambda_)).max() < 1e-12
    assert np.abs(got["lam"] - ev[:7]).max() < 1e-10
    assert fo.subspace_angle(got["q"].astype(complex), np.asarray(r1.q, dtype=complex)) < 1e-8
    assert int(got["info_b"]) == 0 and int(got["M_b"]) == 7 and np.abs(got["lam_b"] - ev[:7]).max() < 1e-10
    assert int(got["info_m"]) == 0 and int(got["M_m"]) == 7 and np.abs(got["lam_m"] - ev[:7]).max() < 1e-10
    assert float(got["res_m"]) < 1e-12 and int(got["fp32_sweeps"]) >= 2


def _rows_worker(rank, world, port, out):
    sys.path.insert(0, str(ROOT / "oracle"))
    import torch
    import torch.distributed as dist
    import feast_oracle as fo
    import feast_port as fp
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    N, M0 = 8, 13
    n = N ** 3
    A = fo.laplacian_3d(N).astype(float).tocsr()
    ev = fo.laplacian_3d_eigs(N)
    Emin, Emax = 0.0, 0.5 * (ev[6] + ev[7])
    Q0 = fo.seeded_subspace(n, M0, complex_storage=False)
    bounds = [n * r // world for r in range(world + 1)]
    r0, r1 = bounds[rank], bounds[rank + 1]
    counts = {"exchange": 0, "small": 0, "small_bytes": 0}

    def exchange(X_loc):                    # all-gather of row blocks: the halo exchange of the CUDA design moves a subset of it
        counts["exchange"] += 1
        m = X_loc.shape[1]
        parts = [torch.zeros((bounds[r + 1] - bounds[r], m), dtype=torch.float64) for r in range(world)]
        dist.all_gather(parts, torch.from_numpy(np.ascontiguousarray(X_loc)))
        return np.concatenate([p.numpy() for p in parts], axis=0)

    def allreduce_small(a):
        counts["small"] += 1
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64)).copy())
        counts["small_bytes"] += t.numel() * 8
        dist.all_reduce(t)
        return t.numpy()

    r = fp.feast_hrr_mslanczos_rows(A, Emin, Emax, M0, fo.feastinit(), Q0, r0, r1, exchange, allreduce_small, inner_maxiter=800)
    full_q = exchange(r.q) if r.M else np.zeros((n, 0))
    if rank == 0:
        np.savez(out, lam=np.sort(r.lambda_), M=r.M, info=r.info, loop=r.loop, q=full_q, steps=sum(r.stats["lz_steps"]),
                 small=counts["small"], small_bytes=counts["small_bytes"], exchanges=counts["exchange"])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_world_size_2_row_sharded_lanczos_equals_single_rank(tmp_path):
    """The row-sharded scheme planned for the 8-GPU target: every rank owns half the rows of A and of every
    block vector; halo rows are exchanged for each mat-vec, dot products and Gram matrices are small all-reduces, pass 2 needs no
    reduction.  Must reproduce the single-rank solve: same M, eigenvalues, loop count, subspace; the communication per Lanczos step
    is two m-vector all-reduces in pass 1 and none in pass 2."""
    import torch.multiprocessing as mp
    sys.path.insert(0, str(ROOT / "oracle"))
    import feast_oracle as fo
    import feast_port as fp
    out = str(tmp_path / "rows.npz")
    port = _free_port()
    mp.spawn(_rows_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    N, M0 = 8, 13
    A = fo.laplacian_3d(N).astype(float).tocsr()
    ev = fo.laplacian_3d_eigs(N)
    Emin, Emax = 0.0, 0.5 * (ev[6] + ev[7])
    Q0 = fo.seeded_subspace(N ** 3, M0, complex_storage=False)
    r1 = fp.feast_hrr_mslanczos(A, Emin, Emax, M0, fo.feastinit(), Q0, inner_rel=1e-3, inner_maxiter=800, adaptive=True)
    assert int(got["info"]) == r1.info == 0 and int(got["M"]) == r1.M == 7 and int(got["loop"]) == r1.loop
    assert np.abs(got["lam"] - ev[:7]).max() < 1e-10 and np.abs(got["lam"] - np.sort(r1.lambda_)).max() < 1e-11
    assert fo.subspace_angle(got["q"].astype(complex), np.asarray(r1.q, dtype=complex)) < 1e-8
    steps = int(got["steps"])
    assert abs(steps - sum(r1.stats["lz_steps"])) <= 3
    # communication budget: 2 small all-reduces per pass-1 step (+1 for ||b|| per sweep), a handful per Rayleigh-Ritz stage
    sweeps = int(got["loop"]) + 1
    assert int(got["small"]) <= 2 * steps + 12 * sweeps
    assert int(got["small_bytes"]) <= (2 * steps + 4 * sweeps) * M0 * 8 + 8 * sweeps * M0 * M0 * 8
