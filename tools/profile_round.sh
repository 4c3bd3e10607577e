#!/bin/bash
# One gpurun call (1 GPU, ~6 minutes of box time; every step has its own timeout) that refreshes what profiles/ needs:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/profile_round.sh r2 > gpurun_out/r2_profile_round.log 2>&1; tail -40 gpurun_out/r2_profile_round.log'
# then, HERE (ncu reads reports without a GPU):
#   python tools/summarize_launches.py gpurun_out/${TAG}_launches.csv > profiles/${TAG}_launches_step.txt
#   python tools/ncu_extract.py gpurun_out/${TAG}_cin.ncu-rep --json profiles/${TAG}_cin_tc_traffic.json > profiles/${TAG}_cin_tc_ncu_summary.txt
#   python tools/ncu_extract.py gpurun_out/${TAG}_hbm.ncu-rep > profiles/${TAG}_hbm_kernels_ncu.txt
TAG=${1:-rN}
mkdir -p gpurun_out
# 1. the headline bench line (never under a profiler)
timeout -s KILL 300 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
cut -c1-300 gpurun_out/${TAG}_bench_n1.json
# 2. launch list of one eager step: shares only (the product path replays a CUDA graph of exactly these launches)
DTB_CUDA_GRAPH=0 timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launch_bench.log 2>&1
# 3. full-section capture of the CIN kernels of one training forward + backward (auto precision = fp16 single pass)
REPS=1 timeout -s KILL 600 ncu --set full --clock-control none --import-source on \
    -k regex:"cin_tc2_fwd_kernel|cin_tc2_dgrad_kernel|cin_tc2_wgrad_kernel" -c 5 -f -o gpurun_out/${TAG}_cin python tools/cin_once.py > gpurun_out/${TAG}_ncu_cin.log 2>&1
# 4. counters of the bandwidth-bound kernels and the Dense GEMMs (one launch each)
timeout -s KILL 600 ncu --set full --clock-control none -k regex:"fm_linear|concat_|col_reduce|bn_apply|bn_bwd_apply|cross_fwd|cross_bwd|adam_rows|dense_tc_rows|dense_tc_wgrad" \
    -c 24 -f -o gpurun_out/${TAG}_hbm python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/${TAG}_ncu_hbm.log 2>&1
# 5. un-profiled timings: CIN kernels (fp16 single pass and bf16x3), bandwidth-bound kernels alone
REPS=3 CHECKB=1 PREC=4 timeout -s KILL 90 python tools/cin_once.py 2>&1 | tail -3
REPS=3 PREC=2 timeout -s KILL 90 python tools/cin_once.py 2>&1 | tail -1
timeout -s KILL 120 python tools/bench_hbm.py > gpurun_out/${TAG}_hbm_kernels.txt 2>&1
tail -12 gpurun_out/${TAG}_hbm_kernels.txt
