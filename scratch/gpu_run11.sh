cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_dense_band.py -m gpu -q -x 2>&1 | tail -40
