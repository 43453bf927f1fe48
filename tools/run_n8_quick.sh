tag=$1
mkdir -p gpurun_out
run() { name=$1; np=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $np "$@" > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_${name}.json').read().strip().splitlines()[-1])
    print('$name', 'value', round(d['value'],3), 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],3), 'int_issue', round(d['int_issue']['frac'],3))
    print('   ', {k: round(v,2) for k,v in d['kernel_ms_per_step_rank0'].items()}, d['wall_ms_rank0'])
    print('   ', d.get('parity'))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/${tag}_${name}.err').read()[-1500:])
PY
}
run c1_n8 8 --steps 2 --warmup 1 --no-cpu-baseline
run c4_n8 8 --config 4 --steps 2 --warmup 1 --no-cpu-baseline
