cd $GRAFT_REPO_ROOT
FEASTCUDA_VERBOSE=1 timeout 900 python -m pytest tests/test_gpu_general.py -m gpu -q -x 2>&1 | tail -40
