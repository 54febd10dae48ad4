cd $GRAFT_REPO_ROOT
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/r10_bench2.json 2> gpurun_out/r10_bench2.err; echo rc=$?; cat gpurun_out/r10_bench2.json | cut -c1-1500; tail -5 gpurun_out/r10_bench2.err
python -c "import __graft_entry__ as g; g.smoke()"
