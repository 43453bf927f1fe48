mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "(sketch or search or h1024 or filter) and not full_parity" ) > gpurun_out/r2n_pytest.log 2>&1; tail -3 gpurun_out/r2n_pytest.log
for v in "MHAPB_K1B_PASSES=1" "MHAPB_K1B_PASSES=0"; do
  echo "== $v"
  env $v timeout 400 python bench.py --config 4 --reads 60000 --no-cpu-baseline --steps 2 --warmup 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step_rank0']; print('minhash',round(k['minhash_ms'],1),'dedup',round(k['hash_dedup_ms'],1),'ms/step',round(d['ms_per_step'],1),'int_issue',round(d['int_issue']['frac'],3), d['parity']['digest'], d['counters'])" | tee -a gpurun_out/r2n_h1024.log
done
timeout 300 python bench.py --no-cpu-baseline --steps 3 --warmup 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('config1 ms/step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],d['kernel_ms_per_step_rank0'],d['parity']['digest'])"
