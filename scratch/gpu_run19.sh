cd $GRAFT_REPO_ROOT
python scratch/probe_msl.py 100 64 1e-3 48 1 0 1 > gpurun_out/r19_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_lz_spmm -s 20 -c 1 -o gpurun_out/r19_lz_p1 python scratch/probe_msl.py 100 64 1e-3 48 1 0 1 > gpurun_out/r19_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lz_spmm -s 60 -c 1 -o gpurun_out/r19_lz_p2 python scratch/probe_msl.py 100 64 1e-3 48 1 0 1 > gpurun_out/r19_ncu2.log 2>&1
ncu --set full --clock-control none -k regex:k_lz_update -s 20 -c 1 -o gpurun_out/r19_lz_upd python scratch/probe_msl.py 100 64 1e-3 48 1 0 1 > gpurun_out/r19_ncu3.log 2>&1
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo rc=$?
python -m pytest tests -m gpu -q 2>&1 | tail -5
