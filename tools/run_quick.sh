# quick check: search parity tests + one bench line (kernel breakdown); usage: bash tools/run_quick.sh <tag> [pytest -k expr]
tag=${1:-rq}; kexpr=${2:-"search or sketch"}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "$kexpr and not full_parity" ) > gpurun_out/${tag}_pytest.log 2>&1
tail -4 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -3 gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','kernel_ms_per_step_rank0','wall_ms_rank0','e2e','gpu_launches','parity','int_issue'):
        print(k, d.get(k))
except Exception as e:
    print('bench parse failed', e)
PY
