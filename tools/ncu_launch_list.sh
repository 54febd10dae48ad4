#!/bin/bash
# Launch list of one converged solve of config 3 (every kernel with its device time) for profiles/, per B200_PROFILING.md:
#   gpurun --timeout 900 -- 'bash tools/ncu_launch_list.sh'
# Runs the same command without ncu first (it must exit 0), then under `ncu --metrics gpu__time_duration.sum`.
# Window: the warm-up solve issues ~9.1 k launches; -s 11500 -c 2500 covers the end of pass 1 (P1, scal1, update, scal2), all of
# pass 2 (LZ_P2_SKIP / LZ_P2_PAIR alternating) and the Rayleigh-Ritz stage of the first timed sweep.  The first attempt of round 1
# died on a 100 s timeout: the profiled process needs ~3 minutes.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
set -o pipefail
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu --no-mixed > gpurun_out/launch_plain.json 2> gpurun_out/launch_plain.err || { echo "plain run failed"; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 11500 -c 2500 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu --no-mixed > gpurun_out/launch_ncu.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/launches.csv
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches.csv")) if len(r) > 5 and r[0].isdigit()]
tot = collections.Counter(); cnt = collections.Counter()
for r in rows:
    name = r[4].split("(")[0]; val = float(r[-1].replace(",", "")); unit = r[-2]
    us = val * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(unit, 1.0)
    tot[name] += us; cnt[name] += 1
s = sum(tot.values())
for k, v in tot.most_common(12):
    print(f"{v / s:6.1%} {v / cnt[k]:9.1f} us x {cnt[k]:5d}  {k[:90]}")
PY
