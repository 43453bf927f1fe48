# one rank's K2b/K2c load of an 8-rank job on one GPU: timings, then ncu --set full of the probe; usage bash tools/run_rank_view.sh <tag>
tag=${1:-rv}
mkdir -p gpurun_out
timeout 60 python tools/probe_rank_view.py --store 100000 --queries 200000 > gpurun_out/${tag}_rank_view.txt 2>&1
cat gpurun_out/${tag}_rank_view.txt | tail -4
timeout 80 ncu --set full --clock-control none --import-source on -k regex:k_probe -c 1 -f -o gpurun_out/${tag}_probe python tools/probe_rank_view.py --store 50000 --queries 100000 --repeat 1 > gpurun_out/${tag}_ncu_probe.log 2>&1
tail -3 gpurun_out/${tag}_ncu_probe.log | cut -c1-300
ls -la gpurun_out | grep ${tag}
