# K2c tunables on the query-mode shape (query sketches in their own block, one candidate per query row: what a rank sees at N=8)
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q -k "search and not full_parity" ) > gpurun_out/r2j_pytest.log 2>&1; tail -3 gpurun_out/r2j_pytest.log
for v in "MHAPB_K2C_CTAS=0 MHAPB_K2C_PREFETCH=1" "MHAPB_K2C_CTAS=0 MHAPB_K2C_PREFETCH=0" "MHAPB_K2C_CTAS=4 MHAPB_K2C_PREFETCH=1" "MHAPB_K2C_CTAS=4 MHAPB_K2C_PREFETCH=0" "MHAPB_K2C_CTAS=2 MHAPB_K2C_PREFETCH=1" "MHAPB_K2C_CTAS=6 MHAPB_K2C_PREFETCH=0"; do
  echo "== $v"
  env $v timeout 300 python bench.py --config 3 --reads 100000 --query-reads 300000 --read-len 10000 --no-cpu-baseline --no-parity --steps 2 --warmup 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step_rank0']; print('probe',round(k['probe_ms'],2),'filter',round(k['filter_ms'],2),'index',round(k['index_ms'],2),'ms/step',round(d['ms_per_step'],1), d['query_counters']['fully_compared'], d['counters']['fully_compared'])" | tee -a gpurun_out/r2j_sweep.log
done
