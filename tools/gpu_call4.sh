#!/bin/bash
# round-2 GPU call 4: fp16 single-pass CIN backward (cin_tc2 dgrad + fp16 wgrad) correctness + timing, full suite,
# ncu --set full of the cin_tc2 kernels
O=gpurun_out/r2c4; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -rfE --timeout 600 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 300 python -m pytest tests/test_zz_baseline_configs_gpu.py -m gpu -q -s -k fp16 > $O/pytest_fp16.log 2>&1
PREC=4 CHECKF=1 CHECKB=1 REPS=3 timeout 300 python tools/cin_once.py > $O/cin_once_p4.log 2>&1
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --cin-precision 4 > $O/bench_p4.json 2> $O/bench_p4.err
PREC=4 REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:cin_tc -c 6 -o $O/cin_p4 python tools/cin_once.py > $O/ncu_cin.log 2>&1
grep -E "passed|failed" $O/pytest.log | tail -n 3; grep -E "^FAILED|^ERROR" $O/pytest.log | head -n 20
grep -E "fp16x1|passed|failed" $O/pytest_fp16.log | tail -n 40
tail -n 5 $O/cin_once_p4.log; cut -c1-300 $O/bench_p4.json; tail -n 3 $O/bench_p4.err; tail -n 3 $O/ncu_cin.log
