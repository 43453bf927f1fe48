tag=$1
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_search.py -m gpu -x -q ) 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline --no-parity 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=2 ms/step',d['ms_per_step'],{k: round(v,2) for k,v in d['kernel_ms_per_step_rank0'].items()})"
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-parity 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1 ms/step',d['ms_per_step'],{k: round(v,2) for k,v in d['kernel_ms_per_step_rank0'].items()})"
