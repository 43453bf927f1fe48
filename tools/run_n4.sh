tag=$1
mkdir -p gpurun_out
run() { name=$1; np=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $np "$@" > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_${name}.json').read().strip().splitlines()[-1])
    print('$name', 'value', round(d['value'],3), 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],3), round(d['e2e']['ms_per_step'],1))
    print('   ', {k: round(v,2) for k,v in d['kernel_ms_per_step_rank0'].items()})
    p=d.get('parity') or {}
    print('   ', {k: p.get(k) for k in ('hits','digest','legs_agree','counters_e2e_equal','single_gpu_equal','oracle_sample_equal')})
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/${tag}_${name}.err').read()[-1500:])
PY
}
run c1_n4 4 --steps 5 --warmup 3 --no-cpu-baseline
CUDA_VISIBLE_DEVICES=0,1 run c1_n2 2 --steps 5 --warmup 3 --no-cpu-baseline
