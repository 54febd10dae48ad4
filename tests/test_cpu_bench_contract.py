"""CPU test of bench.py's output contract: the JSON line of the GPU arm (control flow and post-processing exercised with a stub
engine standing in for libfeastcuda -- no GPU here) and of the reference arm (the CPU port, for real, at a small grid)."""
import json
import sys
import types
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


class StubEngine:
    """Returns fixed timings/statistics where the real engine would run the solve on the device."""

    def __init__(self, fail_mixed=False):
        self.t, self.n, self.fail_mixed = 0.0, None, fail_mixed

    def set_sparse(self, which, A, st):
        self.n = A.shape[0]

    def clear_b(self):
        pass

    def init_distributed(self):
        pass

    def set_row_sharding(self, on=True):
        pass

    def row_range(self):
        return 0, self.n, self.n

    def make_opts(self, **kw):
        return types.SimpleNamespace(**kw)

    def upload_subspace(self, m0, Q):
        pass

    def reset_stats(self):
        self.t = 0.0

    def stats(self):
        k, n, b = [0.0] * 8, [0] * 8, [0.0] * 8
        for i, (ms, by) in {1: (0.48, 1.6e9), 2: (0.24, 1.5e9), 3: (0.55, 2.1e9), 4: (0.47, 1.6e9)}.items():
            k[i], n[i], b[i] = ms * 10, 10, by
        return {"ms_dev_run": self.t, "n_kern": n, "ms_kern": k, "bytes_kern": b, "lz_steps_p1": 3570, "lz_steps_p2": 3570,
                "lz_steps_fp32": 3942, "kernel_launches": 18000}

    def run_interval(self, Emin, Emax, m0, fpm, Z, W, opts):
        if getattr(opts, "mixed", False):
            if self.fail_mixed:
                raise RuntimeError("boom")
            self.t += 1400.0
        else:
            self.t += 2360.0
        return 35, 0, 2e-13, 2

    def solve_interval(self, *a, **kw):
        import feastcuda as fc
        return fc.FeastResult(np.zeros(35), np.zeros((self.n, 35)), 35, np.zeros(35), 0, 2e-13, 2, {})

    def fetch_results(self, m0, M, real):
        return np.zeros(M), np.zeros((self.n, M)), np.full(M, 1e-13)


def _run_bench(monkeypatch, capsys, argv, engine):
    import torch
    import feastcuda as fc
    sys.path.insert(0, str(ROOT))
    import bench
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    monkeypatch.setattr(fc, "default_engine", lambda *a, **k: engine)
    monkeypatch.setattr(sys, "argv", ["bench.py"] + argv)
    bench.main()
    lines = [l for l in capsys.readouterr().out.splitlines() if l.startswith("{")]
    assert len(lines) == 1                                                     # exactly ONE JSON line
    return json.loads(lines[0])


def _check_common(d):
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "e2e", "cpu_baseline"):
        assert key in d, key
    assert d["unit"] == "eigenpairs/s" and d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"].startswith("f64")
    assert "workload" in d["config"] and "model" not in d["config"] and d["data"] == "synthetic"
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])


def test_gpu_arm_json_contract(monkeypatch, capsys):
    d = _run_bench(monkeypatch, capsys, ["--grid", "12", "--m0", "40", "--steps", "2", "--warmup", "3", "--cpu-sample-steps", "2", "--cpu-pair-grid", "8"], StubEngine())
    _check_common(d)
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["scaling"] == "strong"
    assert d["ms_per_step"] == pytest.approx(1400.0) and d["value"] == pytest.approx(35 / 1.4)      # headline: fpm[42] = 1 (FP32 Krylov vectors)
    assert d["config"]["fpm42"] == 1 and "f32" in d["dtype"]
    assert d["gpu_launches"] > 0 and set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert set(d["result"]["parity"]) >= {"all_ranks_ok", "same_M_on_all_ranks", "subspace_angle_vs_analytic", "ranks"}
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["frac"] == pytest.approx(r["achieved"] / r["peak"])
    assert r["traffic"] is None or r["traffic"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and c["unit"] == d["unit"] and "sample" in c
    pair = c["reference_path_pair"]          # the oracle's restatement of the reference's serial path, run to completion beside the engine
    assert pair["n"] == 512 and pair["cpu_true_filter"]["info"] == 0 and pair["cpu_true_filter"]["M"] == 10 and pair["cpu_true_filter"]["seconds"] > 0
    assert set(pair["gpu"]) >= {"seconds", "info", "M"}
    m = d["mixed_precision"]                   # the secondary leg: the other precision setting (FP64 vectors here)
    assert m["ms_per_step"] == pytest.approx(2360.0) and m["result"]["M"] == 35 and "kernels" in m


def test_gpu_arm_survives_a_failing_secondary_leg(monkeypatch, capsys):
    d = _run_bench(monkeypatch, capsys, ["--grid", "12", "--m0", "40", "--steps", "1", "--warmup", "3", "--no-cpu", "--fp64"], StubEngine(fail_mixed=True))
    _check_common(d)
    assert d["ms_per_step"] == pytest.approx(2360.0) and d["cpu_baseline"] is None and d["dtype"] == "f64"
    assert d["mixed_precision"] == {"error": "RuntimeError: boom"}


def test_reference_arm_runs_the_cpu_port(monkeypatch, capsys):
    sys.path.insert(0, str(ROOT))
    import bench
    monkeypatch.setattr(sys, "argv", ["bench.py", "--impl", "reference", "--grid", "14", "--m0", "40", "--steps", "1", "--warmup", "0",
                                      "--cpu-sample-steps", "2"])
    bench.main()
    lines = [l for l in capsys.readouterr().out.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    _check_common(d)
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_banded_line_json_contract_with_the_cpu_oracle_run_to_completion(monkeypatch, capsys):
    """bench.py --config 5 (the banded line): workload construction (band storage = the oracle's, analytic eigenvalues), the roofline
    object from the sampled band kernels and the cpu_baseline leg -- the oracle's restatement of the reference's banded route (LAPACK
    band LU per node) run to completion on the same inputs.  The solve itself is replaced by a stand-in that returns the analytic pairs."""
    import torch
    import feastcuda as fc
    import feast_oracle as fo
    sys.path.insert(0, str(ROOT))
    import bench
    import __graft_entry__ as g
    seen = {}

    def fake_sbev(AB, k, Emin, Emax, M0, fpm, Q0=None):
        A = fo.banded_to_full(np.asarray(AB), k, hermitian=False)
        w, V = np.linalg.eigh(A)
        sel = (w >= Emin) & (w <= Emax)
        seen["n"], seen["k"] = AB.shape[1], k
        stats = {"n_kern": [0] * 6 + [1, 4], "ms_kern": [0.0] * 6 + [2.0, 4.0], "bytes_kern": [0.0] * 6 + [1e6, 8e6], "ms_total": 7.0,
                 "kernel_launches": 40, "lz_steps_p1": 0, "cheb_degree": 0}
        return fc.FeastResult(w[sel], V[:, sel], int(sel.sum()), np.full(int(sel.sum()), 1e-14), 0, 1e-14, 3, stats)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(fc, "dfeast_sbev", fake_sbev)
    monkeypatch.setattr(g, "build", lambda *a, **k: None)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--config", "5", "--n", "700"])
    bench.main()
    lines = [l for l in capsys.readouterr().out.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    _check_common(d)
    assert seen == {"n": 700, "k": 7} and d["config"]["n"] == 700 and "banded" in d["metric"]
    assert d["result"]["M"] == d["result"]["expected_M"] > 20 and d["result"]["max_eig_err_vs_analytic"] < 1e-12
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["kernel"].startswith("k_band_solve_lanes") and r["frac"] == pytest.approx(r["achieved"] / r["peak"])
    assert set(r["all_kernels"]) == {"band_lu", "band_solve"} and r["all_kernels"]["band_solve"]["avg_ms"] == pytest.approx(1.0)
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["value"] > 0 and "run to completion" in c["sample"] and c["max_eig_diff_gpu_vs_cpu"] < 1e-10
