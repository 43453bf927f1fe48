mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "(search or cli or filter) and not full_parity" ) > gpurun_out/r2k_pytest.log 2>&1; tail -3 gpurun_out/r2k_pytest.log
for v in "MHAPB_K2C_STAGE=1" "MHAPB_K2C_STAGE=0"; do
  echo "== $v query-shape"
  env $v timeout 300 python bench.py --config 3 --reads 100000 --query-reads 300000 --read-len 10000 --no-cpu-baseline --no-parity --steps 2 --warmup 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step_rank0']; print('probe',round(k['probe_ms'],2),'filter',round(k['filter_ms'],2),'index',round(k['index_ms'],2),'ms/step',round(d['ms_per_step'],1), d['query_counters']['fully_compared'], d['counters']['fully_compared'])" | tee -a gpurun_out/r2k_sweep.log
  echo "== $v config1"
  env $v timeout 300 python bench.py --no-cpu-baseline --steps 2 --warmup 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step_rank0']; print('probe',round(k['probe_ms'],2),'filter',round(k['filter_ms'],2),'index',round(k['index_ms'],2),'ms/step',round(d['ms_per_step'],1), d['parity']['digest'], d['counters'])" | tee -a gpurun_out/r2k_sweep.log
done
