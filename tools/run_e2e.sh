mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q -k "(search or sketch or filter or cli or abi or jni) and not full_parity" ) 2>&1 | tail -3
MHAPB_TRACE=1 timeout 300 python bench.py --no-cpu-baseline --no-parity --steps 4 --warmup 2 2> gpurun_out/r2p_trace.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],d['e2e']['store_add_reads_rank0'], d['wall_ms_rank0'], d['clocks'])"
grep mhapb gpurun_out/r2p_trace.err | tail -6
