#!/bin/bash
O=gpurun_out/r2c15; mkdir -p $O
timeout 600 python -m pytest tests/test_native_gpu.py tests/test_model_gpu.py tests/test_golden.py -m gpu -q -rfE -k "attention or autoint or cross_validation or golden or oracle" > $O/pytest_sel.log 2>&1
for c in dcn6_autoint4x32 five_nets; do
  timeout 500 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_$c.json 2> $O/bench_$c.err
  DTB_CUDA_GRAPH=0 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches_$c.csv python bench.py --config $c --steps 2 --warmup 3 --no-cpu-baseline --no-graph > $O/ncu_$c.log 2>&1
done
grep -E "passed|failed|FAILED" $O/pytest_sel.log | tail -n 8
for f in $O/bench_*.json; do echo $f; cut -c1-300 $f; done; tail -n 3 $O/bench_*.err
