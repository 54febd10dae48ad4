cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_solve.py tests/test_gpu_stages.py -m gpu -q 2>&1 | tail -4
for eg in 8 4 2; do echo "== egrid $eg"; FEASTCUDA_LZ_EGRID=$eg timeout 300 python scratch/probe_msl.py 100 64 1e-3 3000 1 2>&1 | grep -E "kern|rep |ms_lz" | cut -c1-400; done
python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/r16_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 3000 --csv --log-file gpurun_out/r16_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/r16_ncu.log 2>&1
