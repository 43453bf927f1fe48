# N-GPU bench line only (no pytest, no CPU baseline): usage bash tools/run_n8_bench.sh <tag> <ngpus> [extra bench args]
tag=$1; n=$2; shift 2
mkdir -p gpurun_out
NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/${tag}_bench_n$n.json 2> gpurun_out/${tag}_bench_n$n.err
tail -3 gpurun_out/${tag}_bench_n$n.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_bench_n$n.json').read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','kernel_ms_per_step_rank0','kernel_ms_per_step_max_over_ranks','wall_ms_rank0','gpu_launches','parity','hbm_high_water_gb_max_rank','clocks'):
        print(k, d.get(k))
    print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
except Exception as e:
    print('bench parse failed', e)
PY
