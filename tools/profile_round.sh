#!/bin/bash
# One gpurun call that refreshes everything profiles/ needs for a round and runs every prepared experiment
# (1 GPU, ~10 minutes of box time; every step has its own timeout):
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash tools/profile_round.sh r2 > gpurun_out/r2_profile_round.log 2>&1; tail -60 gpurun_out/r2_profile_round.log'
# then, HERE (ncu reads reports without a GPU):
#   python tools/summarize_launches.py gpurun_out/${TAG}_launches.csv 'adam_rows_kernel<0>' > profiles/${TAG}_launches_step.txt
#   python tools/ncu_extract.py gpurun_out/${TAG}_cin.ncu-rep --json profiles/r1_cin_tc_traffic.json > profiles/${TAG}_cin_tc_ncu_summary.txt
TAG=${1:-rN}
mkdir -p gpurun_out
# 1. the headline bench line (never under a profiler)
timeout -s KILL 240 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
cut -c1-300 gpurun_out/${TAG}_bench_n1.json
# 2. launch list of one step: shares only
timeout -s KILL 150 ncu --metrics gpu__time_duration.sum --clock-control none -s 160 -c 260 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launch_bench.log 2>&1
# 3. full-section capture of the five CIN kernels (one training forward + backward), source-correlated
timeout -s KILL 240 ncu --set full --clock-control none --import-source on -k regex:cin_tc_ -c 5 -f \
    -o gpurun_out/${TAG}_cin python tools/cin_once.py > gpurun_out/${TAG}_ncu_cin.log 2>&1
# 4. A/B of the saved-activation formats and the HBM-bound kernels, un-profiled
REPS=3 timeout -s KILL 60 python tools/cin_once.py 2>&1 | tail -3
FULL=1 REPS=3 timeout -s KILL 60 python tools/cin_once.py 2>&1 | tail -3
for e in 1 2 3 4; do DGRAD_EXP=$e REPS=3 timeout -s KILL 60 python tools/cin_once.py 2>&1 | tail -1; done   # dgrad ablations
DGRAD_EXP=5 CHECK=1 REPS=3 timeout -s KILL 90 python tools/cin_once.py 2>&1 | tail -2                         # dC_hi from shared memory
DGRAD_EXP=6 CHECK=1 REPS=3 timeout -s KILL 90 python tools/cin_once.py 2>&1 | tail -2                         # dgrad on a single fp16 pass
DGRAD_EXP=7 CHECK=1 REPS=3 timeout -s KILL 90 python tools/cin_once.py 2>&1 | tail -2                         # dgrad + wgrad on a single fp16 pass
# 5. the fp16 single-pass forward (precision code 4): step time + roofline of the forward kernel, and its parity tests
timeout -s KILL 120 python bench.py --cin-precision 4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_n1_f16fwd.json 2> gpurun_out/${TAG}_bench_n1_f16fwd.err
cut -c1-300 gpurun_out/${TAG}_bench_n1_f16fwd.json
timeout -s KILL 120 python bench.py --cin-precision 4 --cin-exp 7 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_n1_f16all.json 2> gpurun_out/${TAG}_bench_n1_f16all.err
cut -c1-300 gpurun_out/${TAG}_bench_n1_f16all.json
timeout -s KILL 120 python -m pytest tests/test_zz_baseline_configs_gpu.py -q -k fp16 --runxfail 2>&1 | tail -5
timeout -s KILL 90 python tools/bench_hbm.py > gpurun_out/${TAG}_hbm_kernels.txt 2>&1
tail -12 gpurun_out/${TAG}_hbm_kernels.txt
