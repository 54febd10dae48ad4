cd $GRAFT_REPO_ROOT
python scratch/probe_msl.py 100 64 1e-3 48 1 0 1 > gpurun_out/r6_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_lz_spmm -s 20 -c 1 -o gpurun_out/r6_lz_p1 python scratch/probe_msl.py 100 64 1e-3 48 1 0 1 > gpurun_out/r6_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lz_spmm -s 60 -c 1 -o gpurun_out/r6_lz_p2 python scratch/probe_msl.py 100 64 1e-3 48 1 0 1 > gpurun_out/r6_ncu2.log 2>&1
ncu --set full --clock-control none -k regex:k_lz_update -s 20 -c 1 -o gpurun_out/r6_lz_upd python scratch/probe_msl.py 100 64 1e-3 48 1 0 1 > gpurun_out/r6_ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep
