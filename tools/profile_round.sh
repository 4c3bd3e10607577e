#!/bin/bash
# One gpurun call (1 GPU, ~9 minutes of box time; every step has its own timeout) that refreshes what profiles/ needs:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/profile_round.sh r2 > gpurun_out/r2_profile_round.log 2>&1; tail -60 gpurun_out/r2_profile_round.log'
# The ncu reports stay in /tmp on the box (gpurun copies back at most 64 MiB; the first r2 run lost everything to that limit):
# the summaries are extracted there (tools/ncu_extract.py, tools/summarize_launches.py) and only text comes back.
# SKIP_TESTS=1 skips step 0.
TAG=${1:-rN}
R=/tmp/ncu_$TAG; mkdir -p gpurun_out $R
# 0. the GPU suite and the smoke entry point
if [ -z "$SKIP_TESTS" ]; then
timeout -s KILL 600 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -n 3 gpurun_out/${TAG}_pytest_gpu.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
fi
# 1. the bench lines (never under a profiler): headline config with the CPU baseline, the other three without
timeout -s KILL 300 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
cut -c1-300 gpurun_out/${TAG}_bench_n1.json
for c in deepfm_bs8192 dcn6_autoint4x32 five_nets; do
  timeout -s KILL 300 python bench.py --config $c --no-cpu-baseline > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err
  cut -c1-220 gpurun_out/${TAG}_bench_$c.json
done
# 2. launch lists of eager steps: shares only (the product path replays a CUDA graph of exactly these launches)
DTB_CUDA_GRAPH=0 timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
    --log-file $R/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $R/launch_bench.log 2>&1
python tools/summarize_launches.py $R/launches.csv > gpurun_out/${TAG}_launches_step.txt
for c in dcn6_autoint4x32 five_nets; do
  DTB_CUDA_GRAPH=0 timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv \
      --log-file $R/launches_$c.csv python bench.py --config $c --steps 2 --warmup 3 --no-cpu-baseline > $R/launch_bench_$c.log 2>&1
  python tools/summarize_launches.py $R/launches_$c.csv > gpurun_out/${TAG}_launches_step_$c.txt
done
# 3. full-section capture of the CIN kernels of one training forward + backward (auto precision = fp16 single pass)
REPS=1 timeout -s KILL 600 ncu --set full --clock-control none \
    -k regex:"cin_tc2_fwd_kernel|cin_tc2_dgrad_kernel|cin_tc2_wgrad_kernel" -c 5 -f -o $R/cin python tools/cin_once.py > $R/ncu_cin.log 2>&1
python tools/ncu_extract.py $R/cin.ncu-rep --json gpurun_out/${TAG}_cin_tc_traffic.json > gpurun_out/${TAG}_cin_tc_ncu_summary.txt 2>&1
# 4. counters of the bandwidth-bound kernels and the Dense GEMMs (one launch each), and of the PNN / attention kernels
timeout -s KILL 600 ncu --set full --clock-control none -k regex:"fm_linear|concat_|col_reduce|bn_apply|bn_bwd_apply|cross_fwd|cross_bwd|adam_rows|dense_tc_rows|dense_tc_wgrad" \
    -c 24 -f -o $R/hbm python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > $R/ncu_hbm.log 2>&1
python tools/ncu_extract.py $R/hbm.ncu-rep > gpurun_out/${TAG}_hbm_kernels_ncu.txt 2>&1
REPS=1 timeout -s KILL 300 ncu --set full --clock-control none -k regex:"pnn_|attention_core" -c 10 -f -o $R/pnn_att python tools/pnn_once.py > $R/ncu_pnn.log 2>&1
python tools/ncu_extract.py $R/pnn_att.ncu-rep > gpurun_out/${TAG}_pnn_attention_ncu.txt 2>&1
# 5. un-profiled timings: CIN kernels (fp16 single pass and bf16x3), bandwidth-bound kernels, Dense GEMMs, PNN / attention alone
REPS=3 CHECKB=1 PREC=4 timeout -s KILL 90 python tools/cin_once.py 2>&1 | tail -3
REPS=3 PREC=2 timeout -s KILL 90 python tools/cin_once.py 2>&1 | tail -1
timeout -s KILL 120 python tools/bench_hbm.py > gpurun_out/${TAG}_hbm_kernels.txt 2>&1
tail -12 gpurun_out/${TAG}_hbm_kernels.txt
timeout -s KILL 120 python tools/dense_once.py > gpurun_out/${TAG}_dense_kernels.txt 2>&1; cat gpurun_out/${TAG}_dense_kernels.txt
timeout -s KILL 120 python tools/pnn_once.py > gpurun_out/${TAG}_pnn_attention_kernels.txt 2>&1; cat gpurun_out/${TAG}_pnn_attention_kernels.txt
