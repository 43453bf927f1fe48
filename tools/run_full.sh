tag=$1
mkdir -p gpurun_out
# gate: the sketch parity tests first, with a short leash (a hung kernel must not eat the budget)
( timeout 400 python -m pytest tests/test_gpu_sketch.py tests/test_gpu_filter.py -x -q ) > gpurun_out/${tag}_gate.log 2>&1 || { echo "GATE FAILED"; tail -30 gpurun_out/${tag}_gate.log; exit 1; }
tail -2 gpurun_out/${tag}_gate.log
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('quick bench ms/step',d['ms_per_step'],d['kernel_ms_per_step_rank0'])"
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
# ncu --set full of the K2c kernel at configs[1] (VERDICT r1 missing #6)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_filter_warp" -c 1 -f -o gpurun_out/${tag}_k2c python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-parity > gpurun_out/${tag}_ncu_k2c.log 2>&1
ls -la gpurun_out | grep ${tag}
