tag=$1
mkdir -p gpurun_out
( time timeout 1700 python -m pytest tests -m gpu -x -q --durations=6 ) > gpurun_out/${tag}_pytest.log 2>&1
tail -14 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -3 gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','kernel_ms_per_step_rank0','wall_ms_rank0','e2e','gpu_launches','parity','int_issue','cpu_baseline','clocks'):
        print(k, d.get(k))
except Exception as e:
    print('bench parse failed', e)
PY
