"""GPU, >= 2 devices: the sharded solves (NCCL all-reduce of the accumulator) agree with analytic eigenpairs on every rank."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def test_two_rank_sharded_solves(tmp_path, built_lib):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = tmp_path / "mgpu.json"
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(ROOT / "tests" / "mgpu_worker.py"), str(out)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    res = json.loads(out.read_text())
    for name, r in res.items():
        assert r["ok_all_ranks"], (name, r)
        assert r["allreduce_bytes"] > 0 and r["world"] == 2
