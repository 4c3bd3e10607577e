#!/bin/bash
# round-2 GPU call 1: state of the suite without xfail marks + the fp16 single-pass CIN variants (never run on a GPU before)
mkdir -p gpurun_out/r2c1
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out/r2c1
timeout 900 python -m pytest tests -m gpu -q -rfE --timeout 600 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
for e in 0 6 7; do
  DGRAD_EXP=$e CHECK=1 REPS=3 timeout 300 python tools/cin_once.py > $O/cin_once_exp$e.log 2>&1; echo "rc=$?" >> $O/cin_once_exp$e.log
done
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_default.json 2> $O/bench_default.err
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --cin-precision 4 > $O/bench_p4.json 2> $O/bench_p4.err
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --cin-precision 4 --cin-exp 6 > $O/bench_p4e6.json 2> $O/bench_p4e6.err
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --cin-precision 4 --cin-exp 7 > $O/bench_p4e7.json 2> $O/bench_p4e7.err
tail -5 $O/pytest.log; tail -3 $O/cin_once_exp*.log; cat $O/bench_*.json | cut -c1-400
