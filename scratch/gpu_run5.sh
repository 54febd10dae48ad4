cd $GRAFT_REPO_ROOT
python bench.py > gpurun_out/r5_bench.json 2> gpurun_out/r5_bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/r5_bench.json; tail -5 gpurun_out/r5_bench.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r5_ref.json 2>&1; cat gpurun_out/r5_ref.json | cut -c1-400
nproc; free -g | head -2
python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/r5_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 6000 --csv --log-file gpurun_out/r5_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/r5_ncu.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/r5_launches.csv
