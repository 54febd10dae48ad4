cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_solve.py -m gpu -q -k "mslanczos or laplacian3d or ka1 or ka5 or ka11" 2>&1 | tail -3
for cfg in "512 2 64" "1024 1 64"; do set -- $cfg
 echo "== threads $1 ctas $2 tile $3"
 FEASTCUDA_LZ_THREADS=$1 FEASTCUDA_LZ_CTAS=$2 FEASTCUDA_LZ_TILE=$3 timeout 300 python scratch/probe_msl.py 100 64 1e-3 64 1 0 1 2>&1 | grep -E "kern"
done
