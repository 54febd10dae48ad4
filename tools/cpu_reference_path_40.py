#!/usr/bin/env python3
"""The oracle's restatement of the reference's :serial sparse path (oracle/feast_oracle.py:feast_scsrev = _feast_sparse_hermitian,
sparse/feast_sparse.jl:246-499: sequential node loop, one sparse LU per node -- SuperLU for UMFPACK --, Rayleigh-Ritz, refinement; true
filter) run TO COMPLETION at 40^3 (n = 64 000; SURVEY 8d asks for an end-to-end CPU/GPU pair at this size) on the cores of the machine it
is started on.  No GPU involved: the GPU half of the pair is `python bench.py --grid 40 --no-cpu`.

    python tools/cpu_reference_path_40.py [grid]      ->  one JSON line
"""
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
import numpy as np  # noqa: E402
import bench  # noqa: E402
import feast_oracle as fo  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 40
A, ev, Emin, Emax, Q0 = bench.workload(N, 64)
t0 = time.perf_counter()
r = fo.feast_scsrev(A.tocsc(), Emin, Emax, 64, fo.feastinit(), Q0=np.asarray(Q0).astype(complex), solver="direct", filter="true")
sec = time.perf_counter() - t0
print(json.dumps({"what": "CPU restatement of FeastKit's :serial sparse path (direct solver, true filter), run to completion", "grid": N, "n": N ** 3,
                  "M0": 64, "cores": os.cpu_count(), "seconds": sec, "info": int(r.info), "M": int(r.M), "loops": int(r.loop),
                  "epsout": float(r.epsout), "eigenpairs_per_s": r.M / sec,
                  "max_eig_err_vs_analytic": float(np.abs(np.sort(r.lambda_.real) - ev[:r.M]).max()) if r.M else None}))
