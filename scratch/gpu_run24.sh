cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_solve.py -m gpu -q 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu 2>/dev/null | cut -c1-3000 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('N=2 ms',d['ms_per_step'],'steps',d['result']['lanczos_steps_per_solve'],{n:round(v['avg_ms'],4) for n,v in d['roofline']['all_kernels'].items()})
"
timeout 900 python scratch/run_c2.py 8192 2>&1 | tail -3
