cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_solve.py tests/test_gpu_stages.py -m gpu -q 2>&1 | tail -3
for un in 4 8; do echo "== UN $un"; FEASTCUDA_LZ_UN=$un timeout 300 python scratch/probe_msl.py 100 64 1e-3 64 1 0 1 2>&1 | grep -E "kern" ; done
