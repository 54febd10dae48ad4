set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_tests.log; tail -8 gpurun_out/r2_tests.log
python scratch/probe_msl.py 100 64 1e-3 48 1 0 1 > gpurun_out/r2_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_lz_spmm -s 20 -c 2 -o gpurun_out/r2_lz_p1 python scratch/probe_msl.py 100 64 1e-3 48 1 0 1 > gpurun_out/r2_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lz_spmm -s 60 -c 2 -o gpurun_out/r2_lz_p2 python scratch/probe_msl.py 100 64 1e-3 48 1 0 1 > gpurun_out/r2_ncu2.log 2>&1
tail -3 gpurun_out/r2_ncu1.log gpurun_out/r2_ncu2.log
