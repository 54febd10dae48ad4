cd $GRAFT_REPO_ROOT
FEASTCUDA_VERBOSE=1 timeout 900 python -m pytest tests/test_gpu_configs.py -m gpu -q -k config4 2>&1 | tail -30 | cut -c1-200
