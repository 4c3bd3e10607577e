#!/bin/bash
# round-2 GPU call 3: full suite; cin_tc2 forward (two threads per row) timing + check; dgrad experiment builds 1-5
# (what bounds the data-gradient kernel); launch list of one default step
O=gpurun_out/r2c3; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -rfE --timeout 600 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
PREC=4 CHECKF=1 REPS=3 timeout 300 python tools/cin_once.py > $O/cin_once_p4_v2.log 2>&1
PREC=4 V1=1 CHECKF=1 REPS=3 timeout 300 python tools/cin_once.py > $O/cin_once_p4_v1.log 2>&1
for e in 1 2 3 4 5; do
  DGRAD_EXP=$e REPS=2 timeout 300 python tools/cin_once.py > $O/cin_once_exp$e.log 2>&1
done
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --cin-precision 4 > $O/bench_p4.json 2> $O/bench_p4.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1
grep -E "passed|failed" $O/pytest.log | tail -n 3; grep -E "^FAILED" $O/pytest.log | head -n 20
tail -n 2 $O/cin_once_p4_v2.log $O/cin_once_p4_v1.log; tail -q -n 1 $O/cin_once_exp*.log; cut -c1-300 $O/bench_p4.json; tail -n 3 $O/bench_p4.err
