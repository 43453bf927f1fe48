mkdir -p gpurun_out
nproc > gpurun_out/r2a_host.txt; free -g >> gpurun_out/r2a_host.txt; nvidia-smi -L >> gpurun_out/r2a_host.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/r2a_pytest.log 2>&1
tail -15 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
cut -c1-600 gpurun_out/r2a_bench.json
# ncu --set full of the default K2c and K2b kernels at configs[1] (VERDICT missing #6)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_filter_warp|k_probe" -c 2 -f -o gpurun_out/r2a_k2 python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2a_ncu_k2.log 2>&1
ls -la gpurun_out | tail -6
