#!/usr/bin/env python3
"""The oracle's restatement of the reference's sparse Hermitian path (oracle/feast_oracle.py:feast_hrr = _feast_sparse_hermitian,
sparse/feast_sparse.jl:246-499: one sparse LU per node -- SuperLU for UMFPACK --, Rayleigh-Ritz, refinement; true filter) run TO COMPLETION
at 40^3 (n = 64 000; SURVEY 8d asks for an end-to-end CPU/GPU pair at this size) on the cores of the machine it is started on.

    python tools/cpu_reference_path_40.py [grid]             # :serial  -- sequential node loop
    python tools/cpu_reference_path_40.py [grid] --threads   # :threads -- one worker process per quadrature node, each with its cached LU
                                                             #            (Threads.@threads over the nodes, parallel/feast_parallel.jl:586)
No GPU involved: the GPU half of the pair is `python bench.py --grid 40 --no-cpu`.  Prints one JSON line.
"""
import json
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
import numpy as np  # noqa: E402
import bench  # noqa: E402
import feast_oracle as fo  # noqa: E402

_W = {}


def _worker_init(A, z, weight):
    _W.update(A=A, z=z, weight=weight, fac=None)


def _worker_solve(rhs):
    if _W["fac"] is None:
        _W["fac"] = fo._factor(_W["A"], None, _W["z"])
    return _W["weight"] * fo._solve_factor(_W["fac"], rhs)


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    threads = "--threads" in sys.argv
    N = int(args[0]) if args else 40
    A, ev, Emin, Emax, Q0 = bench.workload(N, 64)
    Ac = A.tocsc()            # real symmetric pencil + real basis + true filter: the oracle's real mode (rho = Re g), like feast_scsrev
    fpm = fo.feastinit()
    fo.feastdefault(fpm)
    Zne, Wne = fo.feast_contour(Emin, Emax, fpm)
    pools, node_pool = [], None
    if threads:
        pools = [ProcessPoolExecutor(1, initializer=_worker_init, initargs=(Ac, z, 2 * w)) for z, w in zip(Zne, Wne)]

        def node_pool(rhs):
            futs = [p.submit(_worker_solve, rhs) for p in pools]
            return sum(f.result() for f in futs)
    t0 = time.perf_counter()
    r = fo.feast_hrr(Ac, None, Emin, Emax, 64, fo.feastinit(), Q0=np.asarray(Q0).astype(complex), solver="direct", filter="true", node_pool=node_pool)
    sec = time.perf_counter() - t0
    for p in pools:
        p.shutdown()
    print(json.dumps({"what": "CPU restatement of FeastKit's sparse Hermitian path (direct solver, true filter), run to completion",
                      "backend": ":threads analog (one worker process per quadrature node)" if threads else ":serial (sequential node loop)",
                      "grid": N, "n": N ** 3, "M0": 64, "cores": os.cpu_count(), "seconds": sec, "info": int(r.info), "M": int(r.M), "loops": int(r.loop),
                      "epsout": float(r.epsout), "eigenpairs_per_s": r.M / sec,
                      "max_eig_err_vs_analytic": float(np.abs(np.sort(r.lambda_.real) - ev[:r.M]).max()) if r.M else None}))


if __name__ == "__main__":
    main()
