cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_solve.py -m gpu -q -k "mslanczos or laplacian3d" 2>&1 | tail -3
for m0 in 8 16 32; do echo "== M0 $m0"; timeout 300 python scratch/probe_msl.py 100 $m0 1e-3 64 1 0 1 2>&1 | grep -E "kern" ; done
