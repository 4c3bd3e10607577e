#!/bin/bash
# round-2 GPU call 10: full suite (zero-bound fix, CUDA-graph train step), five-net debug, CIN timings, all four bench configs
O=gpurun_out/r2c10; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -rfEs --timeout 900 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 200 python tools/debug_five.py > $O/debug_five_p0.log 2>&1
PREC=4 CHECKB=1 REPS=3 timeout 300 python tools/cin_once.py > $O/cin_once_p4.log 2>&1
for c in xdeepfm deepfm_bs8192 dcn6_autoint4x32 five_nets; do
  timeout 500 python bench.py --config $c --steps 20 --warmup 5 > $O/bench_$c.json 2> $O/bench_$c.err
done
DTB_CUDA_GRAPH=0 timeout 400 python bench.py --config deepfm_bs8192 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_deepfm_nograph.json 2> $O/bench_deepfm_nograph.err
grep -E "passed|failed" $O/pytest.log | tail -n 2; grep -E "^FAILED|^ERROR" $O/pytest.log | head -n 30
grep -E "^step" $O/debug_five_p0.log; tail -n 3 $O/cin_once_p4.log
for f in $O/bench_*.json; do echo $f; cut -c1-330 $f; done; tail -n 4 $O/bench_*.err
