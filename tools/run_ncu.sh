tag=$1
mkdir -p gpurun_out
# launch list of one bench step (shares only: per-launch times under ncu are cold-cache and serialised)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/${tag}_ncu_bench.log 2>&1
# full captures: K1b on a quarter of the reads (one launch), the other kernels on the full config
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_minhash_bs2 -c 1 -f -o gpurun_out/${tag}_k1b python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-parity --reads 25000 > gpurun_out/${tag}_ncu_k1b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_hash_dedup|k_ordered|k_filter_warp|k_probe|k_index_count|k_index_fill|k_index_pack|k_scan_apply|k_scan_tiles|k_compact" -c 11 -f -o gpurun_out/${tag}_others python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-parity --reads 25000 > gpurun_out/${tag}_ncu_others.log 2>&1
ls -la gpurun_out | grep ${tag}
