cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_solve.py -m gpu -q -k "float32" 2>&1 | tail -3
python scratch/run_c2.py 4096 > gpurun_out/r22_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60000 --csv --log-file gpurun_out/r22_launches.csv python scratch/run_c2.py 4096 > gpurun_out/r22_ncu.log 2>&1
tail -4 gpurun_out/r22_plain.log
python scratch/run_c2.py 4096 > gpurun_out/r22_plain2.log 2>&1 && \
ncu --set full --clock-control none -k regex:k_zgemm_dmma -s 300 -c 1 -o gpurun_out/r22_zgemm python scratch/run_c2.py 4096 > gpurun_out/r22_ncu2.log 2>&1
