cd $GRAFT_REPO_ROOT
for cfg in "1e-1 0" "3e-2 0" "1e-2 0" "3e-3 0" "1e-3 0" "1e-2 1e-1" "1e-3 1e-2" "1e-2 1e-3"; do set -- $cfg
 echo "== inner_rel $1 inner_rel0 $2"
 FEASTCUDA_VERBOSE=1 timeout 300 python scratch/probe_msl.py 100 64 $1 4000 1 $2 30 2>&1 | grep -E "rep |lanczos k|epsout" | cut -c1-150
done
