#!/bin/bash
# Second (short) GPU call of the round's last session: the shared-memory band LU / lane-parallel band solve (FEASTCUDA_BAND_IMPL=3)
# and the two-level blocked dense LU (FEASTCUDA_DENSE_NBO=128), each tested on its own so that a failing variant can fall back to
# the configuration the first call verified (BAND_IMPL=2, DENSE_NBO=32), then the bench lines of both paths.
#   gpurun --timeout 440 -- 'DEADLINE=400 bash tools/second_gpu_call.sh'
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
DEADLINE=${DEADLINE:-400}
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
lim() { local left=$(( DEADLINE - ( $(date +%s) - T0 ) )); if [ $left -lt 20 ]; then echo 0; elif [ $left -lt $1 ]; then echo $left; else echo $1; fi; }

stamp "band tests (BAND_IMPL=3 default)"
L=$(lim 120); [ $L -gt 0 ] && timeout $L python -m pytest tests/test_gpu_dense_band.py -m gpu -q -rf -k "banded or fixture or band_block" > $O/g_band.log 2>&1
RC_BAND=$?; tail -6 $O/g_band.log
if [ $RC_BAND -ne 0 ]; then export FEASTCUDA_BAND_IMPL=2; stamp "BAND IMPL 3 FAILED (rc=$RC_BAND): continuing with FEASTCUDA_BAND_IMPL=2"; fi

stamp "dense tests (DENSE_NBO=128 default)"
L=$(lim 120); [ $L -gt 0 ] && timeout $L python -m pytest tests/test_gpu_dense_band.py -m gpu -q -rf -k "ka1_ka2 or dense_real or dense_reference or dense_block" > $O/g_dense.log 2>&1
RC_DENSE=$?; tail -6 $O/g_dense.log
if [ $RC_DENSE -ne 0 ]; then export FEASTCUDA_DENSE_NBO=32; stamp "DENSE NBO 128 FAILED (rc=$RC_DENSE): continuing with FEASTCUDA_DENSE_NBO=32"; fi

stamp "bench --config 5 (banded line, n = 10^6)"
L=$(lim 120); [ $L -gt 0 ] && timeout $L python bench.py --config 5 --no-cpu > $O/g_bench_band.json 2> $O/g_bench_band.err
echo "rc=$?"; tail -c 1400 $O/g_bench_band.json; tail -3 $O/g_bench_band.err

stamp "bench --config 1 (dense, n = 8192)"
L=$(lim 90); [ $L -gt 0 ] && timeout $L python bench.py --config 1 > $O/g_bench_dense.json 2> $O/g_bench_dense.err
echo "rc=$?"; tail -c 1100 $O/g_bench_dense.json; tail -3 $O/g_bench_dense.err

stamp "general / config / RCI tests on the dense and banded operators"
L=$(lim 150); [ $L -gt 0 ] && timeout $L python -m pytest tests/test_gpu_general.py tests/test_gpu_configs.py::test_config1_dense_householder_similar_reduced \
  tests/test_gpu_solve.py::test_rci_state_machines_drive_a_full_solve -m gpu -q -rf > $O/g_more.log 2>&1
echo "rc=$?"; tail -5 $O/g_more.log

if [ $RC_DENSE -eq 0 ]; then
  stamp "bench --config 1 with DENSE_NBO=256"
  L=$(lim 90); [ $L -gt 0 ] && FEASTCUDA_DENSE_NBO=256 timeout $L python bench.py --config 1 > $O/g_bench_dense_nbo256.json 2> $O/g_bench_dense_nbo256.err
  echo "rc=$?"; python - <<'PY'
import json
for tag in ("g_bench_dense", "g_bench_dense_nbo256"):
    try:
        d = json.loads(open(f"gpurun_out/{tag}.json").read().strip().splitlines()[-1])
        print(tag, "e2e ms", round(d["ms_per_step"], 1), "device ms", round(d["device_ms_total"], 1), "M", d["result"]["M"], "info", d["result"]["info"],
              "epsout", d["result"]["epsout"], "frac", round(d["roofline"]["frac"], 3))
    except Exception as e:
        print(tag, "no line:", e)
PY
fi

if [ $RC_BAND -eq 0 ]; then
  stamp "ncu --set full: band LU and band solve at n = 10^5"
  L=$(lim 120); [ $L -gt 0 ] && timeout $L ncu --set full --clock-control none --import-source on -k "regex:k_band_lu_smem|k_band_solve_lanes" -c 2 -o $O/g_band \
    python bench.py --config 5 --n 99995 --no-cpu --warmup 0 --steps 1 > $O/g_ncu_band.log 2>&1
  echo "ncu rc=$?"
  L=$(lim 40); [ $L -gt 0 ] && timeout $L ncu -i $O/g_band.ncu-rep --page raw --csv > $O/g_band.raw.csv 2>/dev/null
fi
stamp "done"
for f in $O/g_band.ncu-rep; do [ -f $f ] && [ $(stat -c %s $f) -gt 25000000 ] && rm -f $f; done
ls -la $O | grep " g_" | head -30
