#!/bin/bash
# round-2 GPU call 8: decoupled-tile forward, per-field double-buffered dgrad, unrolled scaler in wgrad: correctness + timing
O=gpurun_out/r2c8; mkdir -p $O
timeout 600 python -m pytest tests/test_zz_baseline_configs_gpu.py tests/test_native_gpu.py -m gpu -q -s -k "fp16 or cin" > $O/pytest_sel.log 2>&1
for pr in 0 2; do DTB_CIN_PRECISION=$pr timeout 200 python tools/debug_five.py > $O/debug_five_p$pr.log 2>&1; done
NETS=cin_nets timeout 200 python tools/debug_five.py > $O/debug_five_cin_only.log 2>&1
PREC=4 CHECKF=1 CHECKB=1 REPS=3 timeout 300 python tools/cin_once.py > $O/cin_once_p4.log 2>&1
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench.json 2> $O/bench.err
PREC=4 REPS=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"cin_tc2_fwd_kernel|cin_tc2_dgrad_kernel|cin_tc2_wgrad_kernel" -c 4 -o $O/cin_p4 python tools/cin_once.py > $O/ncu_cin.log 2>&1
grep -E "passed|failed|FAILED" $O/pytest_sel.log | tail -n 10
tail -n 30 $O/debug_five_p0.log; tail -n 6 $O/debug_five_p2.log; tail -n 12 $O/debug_five_cin_only.log; tail -n 5 $O/cin_once_p4.log; cut -c1-250 $O/bench.json; tail -n 3 $O/bench.err; tail -n 2 $O/ncu_cin.log
