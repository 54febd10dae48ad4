cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_dense_band.py tests/test_gpu_general.py tests/test_gpu_configs.py -m gpu -q 2>&1 | tail -6
timeout 900 python scratch/run_c2.py 8192 2>&1 | tail -5
