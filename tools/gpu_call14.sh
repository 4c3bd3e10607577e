#!/bin/bash
O=gpurun_out/r2c14; mkdir -p $O
timeout 300 python tools/debug_capture.py > $O/debug_capture.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q -rfEs --timeout 900 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
for c in dcn6_autoint4x32 five_nets; do
  timeout 500 python bench.py --config $c --steps 20 --warmup 5 > $O/bench_$c.json 2> $O/bench_$c.err
done
cat $O/debug_capture.log | cut -c1-200; grep -E "passed|failed" $O/pytest.log | tail -n 2; grep -E "^FAILED|^ERROR" $O/pytest.log | head -n 30
for f in $O/bench_*.json; do echo $f; cut -c1-330 $f; done; tail -n 4 $O/bench_*.err
