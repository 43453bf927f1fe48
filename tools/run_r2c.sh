mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q -k "not full_parity" ) > gpurun_out/r2c_pytest.log 2>&1
tail -8 gpurun_out/r2c_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_filter_warp|k_index_count|k_index_fill|k_index_pack|k_scan_apply|k_scan_tiles|k_probe" -c 8 -f -o gpurun_out/r2c_k2 python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-parity > gpurun_out/r2c_ncu_k2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_hash_dedup|k_ordered" -c 2 -f -o gpurun_out/r2c_k1ac python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-parity --reads 25000 > gpurun_out/r2c_ncu_k1ac.log 2>&1
ls -la gpurun_out | tail -5
