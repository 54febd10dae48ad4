cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_solve.py -m gpu -q -k "mslanczos or laplacian3d or ka1 or ka5" 2>&1 | tail -5
for thr in 1024 512; do for tile in 32 64 128 256 512 2048; do
 cta=1; if [ $thr = 512 ]; then cta=2; fi
 echo "== threads $thr ctas $cta tile $tile"
 FEASTCUDA_LZ_THREADS=$thr FEASTCUDA_LZ_CTAS=$cta FEASTCUDA_LZ_TILE=$tile timeout 300 python scratch/probe_msl.py 100 64 1e-3 64 1 0 0 2>&1 | grep -E "kern|rep"
done; done
echo "== full solve default"
timeout 300 python scratch/probe_msl.py 100 64 1e-3 3000 2 2>&1 | grep -E "kern|rep |ms_total"
