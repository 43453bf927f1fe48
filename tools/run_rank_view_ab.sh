# A/B of MHAPB_K2B_L2 on one rank's probe load of an 8-rank job; usage bash tools/run_rank_view_ab.sh <tag>
tag=${1:-rv}
mkdir -p gpurun_out
timeout 45 python tools/probe_rank_view.py --store 100000 --queries 200000 --repeat 2 --l2-modes 0,1,2,3,0 > gpurun_out/${tag}_rank_view_l2_ab.txt 2>&1
tail -12 gpurun_out/${tag}_rank_view_l2_ab.txt | cut -c1-260
