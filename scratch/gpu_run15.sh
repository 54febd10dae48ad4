cd $GRAFT_REPO_ROOT
export FEASTCUDA_LZ_STAGED=1
timeout 900 python -m pytest tests/test_gpu_solve.py -m gpu -q -x -k "mslanczos or laplacian3d or ka1 or ka5 or ka11" 2>&1 | tail -5
timeout 300 python scratch/probe_msl.py 100 64 1e-3 64 1 0 1 2>&1 | grep -E "kern|rep "
FEASTCUDA_LZ_STAGED=0 timeout 300 python scratch/probe_msl.py 100 64 1e-3 64 1 0 1 2>&1 | grep -E "kern|rep "
