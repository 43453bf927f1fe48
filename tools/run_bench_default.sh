# what the driver runs at round end: python bench.py (no flags), then the reference arm; usage bash tools/run_bench_default.sh <tag>
tag=${1:-rd}
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "rc=$?"; tail -2 gpurun_out/${tag}_bench.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','kernel_ms_per_step_rank0','e2e','gpu_launches','parity','int_issue','roofline','cpu_baseline','clocks'):
        print(k, d.get(k))
except Exception as e:
    print('bench parse failed', e)
PY
