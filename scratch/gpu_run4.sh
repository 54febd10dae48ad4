cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_solve.py -m gpu -q -k "mslanczos or laplacian3d or ka1 or ka5" 2>&1 | tail -3
for thr in 1024 512; do for pf in -1 32 256; do for un in 4 8; do
 cta=1; if [ $thr = 512 ]; then cta=2; fi
 echo "== threads $thr ctas $cta pf $pf un $un"
 FEASTCUDA_LZ_THREADS=$thr FEASTCUDA_LZ_CTAS=$cta FEASTCUDA_LZ_PF=$pf FEASTCUDA_LZ_UN=$un timeout 300 python scratch/probe_msl.py 100 64 1e-3 64 1 0 1 2>&1 | grep -E "kern"
done; done; done
