#!/bin/bash
# round-2 GPU call 5: same-activation gradient check of the fp16 backward; ncu --set full of cin_tc2 fwd/dgrad + fp16 wgrad;
# launch list of one step (new dense_tc producers)
O=gpurun_out/r2c5; mkdir -p $O
timeout 600 python -m pytest tests/test_zz_baseline_configs_gpu.py tests/test_native_gpu.py -m gpu -q -s -k "fp16 or dense" > $O/pytest_sel.log 2>&1
PREC=4 CHECKB=1 REPS=2 timeout 300 python tools/cin_once.py > $O/cin_once_p4.log 2>&1
PREC=4 REPS=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"cin_tc2_fwd_kernel|cin_tc2_dgrad_kernel|cin_tc_wgrad_kernel" -c 3 -o $O/cin_p4 python tools/cin_once.py > $O/ncu_cin.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_p4.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --cin-precision 4 > $O/ncu_bench.log 2>&1
grep -E "fp16x1|passed|failed|FAILED" $O/pytest_sel.log | tail -n 40
tail -n 4 $O/cin_once_p4.log; tail -n 3 $O/ncu_cin.log
