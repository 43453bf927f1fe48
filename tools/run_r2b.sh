mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q -k "not full_parity" ) > gpurun_out/r2b_pytest.log 2>&1
tail -15 gpurun_out/r2b_pytest.log
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
tail -5 gpurun_out/r2b_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2b_bench.json').read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','ms_per_step_cuda_events','kernel_ms_per_step_rank0','wall_ms_rank0','counters','e2e','gpu_launches','parity','int_issue'):
        print(k, d.get(k))
except Exception as e:
    print('bench parse failed', e)
PY
