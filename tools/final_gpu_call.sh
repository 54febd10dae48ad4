#!/bin/bash
# The round's closing GPU call (one box, ~12 minutes): the banded kernels in their own process first (a faulting kernel poisons the CUDA
# context of its process), then the whole GPU suite, the bench lines, the A/B run of the round-1 band kernels and two ncu captures.
#   gpurun --timeout 840 -- 'bash tools/final_gpu_call.sh'
# Every step has its own timeout; a step that fails does not stop the later ones.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
DEADLINE=${DEADLINE:-780}        # seconds of box time this script may use in total; later steps shrink or are skipped
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
# lim <seconds>: the step's timeout, capped by what is left before the deadline (0 = skip)
lim() { local left=$(( DEADLINE - ( $(date +%s) - T0 ) )); if [ $left -lt 25 ]; then echo 0; elif [ $left -lt $1 ]; then echo $left; else echo $1; fi; }

stamp "band tests, new kernels"
L=$(lim 240); [ $L -gt 0 ] && timeout $L python -m pytest tests/test_gpu_dense_band.py -m gpu -q > $O/f_band_new.log 2>&1
RC_BAND=$?
tail -4 $O/f_band_new.log
if [ $RC_BAND -ne 0 ]; then
  export FEASTCUDA_BAND_IMPL=1
  stamp "BAND TESTS FAILED with the new kernels (rc=$RC_BAND): the rest of this call runs FEASTCUDA_BAND_IMPL=1"
fi

stamp "full GPU suite"
L=$(lim 480); [ $L -gt 0 ] && timeout $L python -m pytest tests -m gpu -q > $O/f_tests.log 2>&1
echo "tests rc=$?"; tail -4 $O/f_tests.log

stamp "bench (default: configs[2], N=1)"
L=$(lim 300); [ $L -gt 0 ] && timeout $L python bench.py > $O/f_bench_n1.json 2> $O/f_bench_n1.err
echo "bench rc=$?"; tail -c 1500 $O/f_bench_n1.json

stamp "bench --config 5 (banded line)"
L=$(lim 200); [ $L -gt 0 ] && timeout $L python bench.py --config 5 --no-cpu > $O/f_bench_band.json 2> $O/f_bench_band.err
echo "band bench rc=$?"; tail -c 2500 $O/f_bench_band.json; tail -3 $O/f_bench_band.err

stamp "bench --config 1 (dense)"
L=$(lim 200); [ $L -gt 0 ] && timeout $L python bench.py --config 1 > $O/f_bench_dense.json 2> $O/f_bench_dense.err
echo "dense bench rc=$?"; tail -c 1800 $O/f_bench_dense.json; tail -3 $O/f_bench_dense.err

if [ $RC_BAND -eq 0 ]; then
  stamp "band tests, round-1 kernels (FEASTCUDA_BAND_IMPL=1)"
  L=$(lim 200); [ $L -gt 0 ] && FEASTCUDA_BAND_IMPL=1 timeout $L python -m pytest tests/test_gpu_dense_band.py -m gpu -q -k "band or fixture" > $O/f_band_old.log 2>&1
  echo "rc=$?"; tail -3 $O/f_band_old.log
  stamp "bench --config 5 at n = 10^5 with both band implementations"
  L=$(lim 120); [ $L -gt 0 ] && timeout $L python bench.py --config 5 --n 99995 --no-cpu > $O/f_bench_band_1e5_new.json 2> $O/f_bench_band_1e5_new.err
  L=$(lim 200); [ $L -gt 0 ] && FEASTCUDA_BAND_IMPL=1 timeout $L python bench.py --config 5 --n 99995 --no-cpu > $O/f_bench_band_1e5_old.json 2> $O/f_bench_band_1e5_old.err
  python - <<'PY'
import json
for tag in ("new", "old"):
    try:
        d = json.loads(open(f"gpurun_out/f_bench_band_1e5_{tag}.json").read().strip().splitlines()[-1])
        print(tag, "ms", round(d["ms_per_step"], 1), "device", round(d["device_ms_total"], 1), "M", d["result"]["M"], "info", d["result"]["info"],
              {k: round(v["avg_ms"], 3) for k, v in (d["roofline"] or {}).get("all_kernels", {}).items()})
    except Exception as e:
        print(tag, "no line:", e)
PY
fi

stamp "ncu --set full: k_zgemm_dmma_async (trailing update of the dense LU)"
L=$(lim 200); [ $L -gt 0 ] && timeout $L ncu --set full --clock-control none --import-source on -k regex:k_zgemm_dmma_async -s 60 -c 1 -o $O/f_zgemm \
  python bench.py --config 1 --warmup 0 --steps 1 > $O/f_ncu_zgemm.log 2>&1
echo "ncu rc=$?"
L=$(lim 60); [ $L -gt 0 ] && timeout $L ncu -i $O/f_zgemm.ncu-rep --page raw --csv > $O/f_zgemm.raw.csv 2>/dev/null

if [ $RC_BAND -eq 0 ]; then
  stamp "ncu --set full: band kernels at n = 10^5"
  L=$(lim 200); [ $L -gt 0 ] && timeout $L ncu --set full --clock-control none --import-source on -k regex:k_band_ -c 3 -o $O/f_band \
    python bench.py --config 5 --n 99995 --no-cpu --warmup 0 --steps 1 > $O/f_ncu_band.log 2>&1
  echo "ncu rc=$?"
  L=$(lim 60); [ $L -gt 0 ] && timeout $L ncu -i $O/f_band.ncu-rep --page raw --csv > $O/f_band.raw.csv 2>/dev/null
  stamp "compute-sanitizer memcheck on the stage-level band solves"
  L=$(lim 240); [ $L -gt 0 ] && timeout $L compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_dense_band.py -m gpu -q -k "every_kernel_path" \
    > $O/f_sanitizer.log 2>&1
  echo "sanitizer rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $O/f_sanitizer.log | tail -3
fi
stamp "done"
# gpurun merges at most 64 MiB back: the raw CSVs carry the numbers, an oversized report stays on the box
for f in $O/f_band.ncu-rep $O/f_zgemm.ncu-rep; do
  [ -f $f ] && [ $(stat -c %s $f) -gt 25000000 ] && rm -f $f
done
ls -la $O | grep " f_" | head -40
