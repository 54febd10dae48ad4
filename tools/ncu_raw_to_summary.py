#!/usr/bin/env python3
"""Fold `ncu -i X.ncu-rep --page raw --csv` dumps of the Lanczos kernels into profiles/ncu_summary.json.

    python tools/ncu_raw_to_summary.py KEY=path.raw.csv:KERNEL_SUBSTRING[:ALG_BYTES] ...

One entry per KEY: the FIRST launch in the CSV whose kernel name contains KERNEL_SUBSTRING (all launches of one kind in a capture
agree to < 0.5 %).  The captures of round 2 were taken with
    ncu --set full --clock-control none --import-source on -k regex:k_lz32_ -s <skip> -c 4 -o gpurun_out/r2_lz32_pass2 \
        python bench.py --steps 1 --warmup 1 --no-cpu --no-mixed           (and --fp64 for the FP64 twins)
after the same command had exited 0 without ncu; the .ncu-rep files stay in gpurun_out/ (scratch), the raw CSVs of the FP32 kernels are tracked
under profiles/r2/.
"""
import csv
import json
import pathlib
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}

FIELDS = {
    "duration_ms": "gpu__time_duration.sum",
    "dram_read": "dram__bytes_read.sum",
    "dram_write": "dram__bytes_write.sum",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct",
    "l1_hit_pct": "l1tex__t_sector_hit_rate.pct",
    "l1tex_pct": "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "sm_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "mem_pct": "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
    "regs": "launch__registers_per_thread",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "warp_insts": "smsp__inst_executed.sum",
    "issue_active_pct": "sm__inst_issued.avg.pct_of_peak_sustained_active",
}


def first_launch(path, needle):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        if needle in r[hdr.index("Kernel Name")]:
            out = {"kernel": r[hdr.index("Kernel Name")].strip()}
            for k, m in FIELDS.items():
                if m in hdr:
                    i = hdr.index(m)
                    out[k] = float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)
            out["dram_bytes_per_launch"] = out["dram_read"] + out["dram_write"]
            return out
    raise SystemExit(f"{path}: no launch of a kernel containing {needle!r}")


def main(argv):
    prof = ROOT / "profiles" / "ncu_summary.json"
    summ = json.loads(prof.read_text())
    for spec in argv:
        key, rest = spec.split("=", 1)
        parts = rest.split(":")
        e = first_launch(parts[0], parts[1])
        e["source"] = pathlib.Path(parts[0]).name
        if len(parts) > 2:
            e["algorithmic_bytes_per_launch"] = float(parts[2])
        summ[key] = e
    prof.write_text(json.dumps(summ, indent=1) + "\n")


if __name__ == "__main__":
    main(sys.argv[1:])
