#!/bin/bash
# One GPU-box visit: parity tests, bench (default + A/B env variants), ncu launch list and a full capture of K1b.
# usage: tools/gpu_round.sh <tag> [variants...]   (outputs under gpurun_out/)
tag=${1:-rX}; shift
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest.log 2>&1
tail -3 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_bench.json | cut -c1-400
for v in "$@"; do
  echo "== $v"
  env $v timeout 300 python bench.py --no-cpu-baseline --steps 2 --warmup 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps(d['kernel_ms_per_step_rank0']), d['value'], d['ms_per_step'])" | tee -a gpurun_out/${tag}_variants.log
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_minhash_bs2 -s 1 -c 1 -f -o gpurun_out/${tag}_k_minhash python bench.py --steps 1 --warmup 0 --no-cpu-baseline --reads 25000 > gpurun_out/${tag}_ncu_full.log 2>&1
ls -la gpurun_out | tail -12
