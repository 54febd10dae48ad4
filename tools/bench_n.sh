cd $GRAFT_REPO_ROOT
N=$1
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo rc=$?
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo rc=$?
fi
tail -c 2500 gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
