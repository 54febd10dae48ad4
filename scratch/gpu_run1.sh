set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r1_tests.log; tail -5 gpurun_out/r1_tests.log
FEASTCUDA_VERBOSE=1 timeout 300 python scratch/probe_msl.py 40 64 1e-3 2000 1 > gpurun_out/r1_msl40.log 2>&1; tail -12 gpurun_out/r1_msl40.log
FEASTCUDA_VERBOSE=1 timeout 600 python scratch/probe_msl.py 100 64 1e-3 3000 2 > gpurun_out/r1_msl100.log 2>&1; tail -24 gpurun_out/r1_msl100.log
for cfg in "512 2 0" "512 1 0" "1024 1 150" "512 4 0"; do set -- $cfg
FEASTCUDA_LZ_THREADS=$1 FEASTCUDA_LZ_CTAS=$2 FEASTCUDA_LZ_FARW=$3 timeout 300 python scratch/probe_msl.py 100 64 1e-3 3000 1 > gpurun_out/r1_msl100_$1_$2_$3.log 2>&1; grep -E "rep|kern" gpurun_out/r1_msl100_$1_$2_$3.log
done
