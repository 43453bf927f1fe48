# multi-GPU visit: usage bash tools/run_multi.sh <tag> <ngpus>
tag=$1; n=$2
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_gpus.txt; nvidia-smi topo -m >> gpurun_out/${tag}_gpus.txt 2>&1
( time timeout 1200 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q -s ) > gpurun_out/${tag}_multirank_pytest.log 2>&1
tail -12 gpurun_out/${tag}_multirank_pytest.log | cut -c1-600
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 3 --warmup 2 > gpurun_out/${tag}_bench_n$n.json 2> gpurun_out/${tag}_bench_n$n.err
tail -3 gpurun_out/${tag}_bench_n$n.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_bench_n$n.json').read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','kernel_ms_per_step_rank0','wall_ms_rank0','gpu_launches','parity','nccl_version','hbm_high_water_gb_max_rank','cpu_baseline'):
        print(k, d.get(k))
    print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
except Exception as e:
    print('bench parse failed', e)
PY
