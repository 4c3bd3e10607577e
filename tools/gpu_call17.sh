#!/bin/bash
# round-2 call 17: PNN kernels with chunk loops / prefetch, attention backward single sweep + relu-input mask,
# direct Dense epilogue by default; full GPU suite; ncu of the PNN / attention kernels
O=gpurun_out/r2c17; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
timeout 300 python tools/pnn_once.py > $O/pnn_once.log 2>&1
for c in dcn6_autoint4x32 five_nets; do
  timeout 400 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_$c.json 2> $O/bench_$c.err
done
REPS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"pnn_|attention_core" -c 10 -f -o $O/pnn_att python tools/pnn_once.py > $O/ncu_pnn.log 2>&1
tail -n 6 $O/pytest_gpu.log; cat $O/pnn_once.log; for f in $O/bench_*.json; do cut -c1-200 $f; done; tail -n 3 $O/ncu_pnn.log
