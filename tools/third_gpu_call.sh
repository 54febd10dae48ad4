#!/bin/bash
# Third (last, ~3 minute) GPU call: the dense LU with 256-column outer blocks under the dense tests, an ncu capture of one rank-128
# trailing update (the 7th GEMM launch of the first outer block: 3 in-block updates, 3 U12 updates, then the trailing matrix), and the
# dense bench line at 256 / 512.
#   gpurun --timeout 240 -- 'DEADLINE=200 bash tools/third_gpu_call.sh'
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
DEADLINE=${DEADLINE:-200}
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
lim() { local left=$(( DEADLINE - ( $(date +%s) - T0 ) )); if [ $left -lt 15 ]; then echo 0; elif [ $left -lt $1 ]; then echo $left; else echo $1; fi; }
stamp "dense tests with FEASTCUDA_DENSE_NBO=256"
L=$(lim 60); [ $L -gt 0 ] && FEASTCUDA_DENSE_NBO=256 timeout $L python -m pytest tests/test_gpu_dense_band.py tests/test_gpu_configs.py::test_config1_dense_householder_similar_reduced \
  tests/test_gpu_general.py -m gpu -q -rf -k "ka1_ka2 or dense_real or dense_reference or dense_block or config1 or dense or ka4 or generalized_dense" > $O/h_dense256.log 2>&1
echo "rc=$?"; tail -4 $O/h_dense256.log
stamp "ncu --set full: rank-128 trailing update"
L=$(lim 90); [ $L -gt 0 ] && timeout $L ncu --set full --clock-control none --import-source on -k regex:k_zgemm_dmma_async -s 6 -c 1 -o $O/h_zgemm128 \
  python bench.py --config 1 --warmup 0 --steps 1 > $O/h_ncu_zgemm128.log 2>&1
echo "ncu rc=$?"
L=$(lim 30); [ $L -gt 0 ] && timeout $L ncu -i $O/h_zgemm128.ncu-rep --page raw --csv > $O/h_zgemm128.raw.csv 2>/dev/null
for nbo in 256 512; do
  stamp "bench --config 1 with DENSE_NBO=$nbo"
  L=$(lim 60); [ $L -gt 0 ] && FEASTCUDA_DENSE_NBO=$nbo timeout $L python bench.py --config 1 > $O/h_bench_dense_nbo$nbo.json 2> $O/h_bench_dense_nbo$nbo.err
done
python - <<'PY'
import json
for tag in ("h_bench_dense_nbo256", "h_bench_dense_nbo512"):
    try:
        d = json.loads(open(f"gpurun_out/{tag}.json").read().strip().splitlines()[-1])
        print(tag, "e2e ms", round(d["ms_per_step"], 1), "device ms", round(d["device_ms_total"], 1), "M", d["result"]["M"], "info", d["result"]["info"],
              "epsout", d["result"]["epsout"], "eig err", d["result"]["max_eig_err_vs_analytic"], "frac", round(d["roofline"]["frac"], 3))
    except Exception as e:
        print(tag, "no line:", e)
PY
stamp "done"
