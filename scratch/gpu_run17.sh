cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_configs.py -m gpu -q 2>&1 | tail -15
for i in 1 2; do FEASTCUDA_VERBOSE=1 timeout 300 python scratch/probe_msl.py 100 64 1e-3 3000 1 2>&1 | grep -E "rep |lanczos k|epsout=" | cut -c1-200; done
timeout 900 python scratch/run_c2.py 8192 2>&1 | tail -8
