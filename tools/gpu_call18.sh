#!/bin/bash
# round-2 call 18: attention kernels with cp.async double buffering, PNN dK batched gather / dE prefetch
O=gpurun_out/r2c18; mkdir -p $O
timeout 600 python -m pytest tests/test_native_gpu.py tests/test_model_gpu.py tests/test_zz_baseline_configs_gpu.py -m gpu -x -q -k "pnn or attention or autoint or five or config" > $O/pytest_sel.log 2>&1; echo "rc=$?" >> $O/pytest_sel.log
timeout 300 python tools/pnn_once.py > $O/pnn_once.log 2>&1
for c in dcn6_autoint4x32 five_nets; do
  timeout 400 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_$c.json 2> $O/bench_$c.err
done
tail -n 4 $O/pytest_sel.log; cat $O/pnn_once.log; for f in $O/bench_*.json; do cut -c1-200 $f; done
