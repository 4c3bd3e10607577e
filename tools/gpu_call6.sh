#!/bin/bash
# round-2 GPU call 6: cin_tc2 with two MMA issuers, single-sweep dgrad, scaler-warp wgrad: correctness + timing
O=gpurun_out/r2c6; mkdir -p $O
timeout 600 python -m pytest tests/test_zz_baseline_configs_gpu.py tests/test_native_gpu.py -m gpu -q -s -k "fp16 or cin" > $O/pytest_sel.log 2>&1
PREC=4 CHECKF=1 CHECKB=1 REPS=3 timeout 300 python tools/cin_once.py > $O/cin_once_p4.log 2>&1
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --cin-precision 4 > $O/bench_p4.json 2> $O/bench_p4.err
PREC=4 REPS=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"cin_tc2_fwd_kernel|cin_tc2_dgrad_kernel|cin_tc2_wgrad_kernel" -c 4 -o $O/cin_p4 python tools/cin_once.py > $O/ncu_cin.log 2>&1
grep -E "fp16x1|passed|failed|FAILED" $O/pytest_sel.log | tail -n 30
tail -n 5 $O/cin_once_p4.log; cut -c1-250 $O/bench_p4.json; tail -n 3 $O/bench_p4.err; tail -n 2 $O/ncu_cin.log
