# at-size multi-GPU visit: usage bash tools/run_atsize.sh <tag> <ngpus> [small]
# runs: multirank parity (torchrun test only), bench config 1 at N, config 4 at N, config 3 at N/2 (4 of 8 GPUs)
tag=$1; n=$2; small=$3
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_gpus.txt; nvidia-smi topo -m >> gpurun_out/${tag}_gpus.txt 2>&1; nproc >> gpurun_out/${tag}_gpus.txt; free -g >> gpurun_out/${tag}_gpus.txt
run() { # name, nproc, extra args...
  name=$1; np=$2; shift 2
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $np "$@" ) > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  tail -4 gpurun_out/${tag}_${name}.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_${name}.json').read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','kernel_ms_per_step_rank0','wall_ms_rank0','e2e','counters','query_counters','parity','hbm_high_water_gb_max_rank','int_issue','clocks'):
        v=d.get(k)
        if k=='int_issue' and v: v=v.get('frac')
        if k=='e2e' and v: v=(v.get('value'), v.get('ms_per_step'))
        print(' ',k, v)
except Exception as e:
    print('bench parse failed', e)
PY
}
( time timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q -s -k "ranks_reproduce" ) > gpurun_out/${tag}_multirank_pytest.log 2>&1
tail -6 gpurun_out/${tag}_multirank_pytest.log | cut -c1-600
if [ -n "$small" ]; then
  run c1_n$n $n --steps 2 --warmup 1 --no-cpu-baseline --reads 20000
  run c4_n$n $n --config 4 --steps 2 --warmup 1 --no-cpu-baseline --reads 6000
  h=$((n/2)); [ $h -lt 1 ] && h=1
  CUDA_VISIBLE_DEVICES=$(seq -s, 0 $((h-1))) run c3_n$h $h --config 3 --steps 2 --warmup 1 --no-cpu-baseline --reads 8000 --query-reads 8000
else
  run c1_n$n $n --steps 5 --warmup 3
  run c4_n$n $n --config 4 --steps 2 --warmup 1 --no-cpu-baseline
  h=$((n/2))
  CUDA_VISIBLE_DEVICES=$(seq -s, 0 $((h-1))) run c3_n$h $h --config 3 --steps 2 --warmup 1 --no-cpu-baseline
fi
ls -la gpurun_out | grep ${tag}
