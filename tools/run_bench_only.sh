tag=$1
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','ms_per_step_cuda_events','kernel_ms_per_step_rank0','wall_ms_rank0','gpu_launches','parity','cpu_baseline','clocks','roofline'):
    print(k, d.get(k))
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'int_issue', d['int_issue']['frac'])
PY
timeout 300 python bench.py --config 2 --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/${tag}_bench_c2.json 2>/dev/null
python -c "import json; d=json.loads(open('gpurun_out/${tag}_bench_c2.json').read().strip().splitlines()[-1]); print('config2', d['value'], d['ms_per_step'], d['e2e']['value'], d['int_issue']['frac'], d['parity']['digest'], d['kernel_ms_per_step_rank0'])"
