#!/bin/bash
# round-2 call 16: templated PNN kernels + vectorised attention kernels + Dense epilogue variant
O=gpurun_out/r2c16; mkdir -p $O
timeout 600 python -m pytest tests/test_native_gpu.py -m gpu -x -q -k "pnn or attention or dense" > $O/pytest_sel.log 2>&1; echo "rc=$?" >> $O/pytest_sel.log
DTB_DENSE_DIRECT=1 timeout 600 python -m pytest tests/test_native_gpu.py -m gpu -x -q -k "dense" > $O/pytest_dense_direct.log 2>&1; echo "rc=$?" >> $O/pytest_dense_direct.log
timeout 300 python tools/pnn_once.py > $O/pnn_once.log 2>&1
timeout 300 python tools/dense_once.py > $O/dense_once_0.log 2>&1
DTB_DENSE_DIRECT=1 timeout 300 python tools/dense_once.py > $O/dense_once_1.log 2>&1
for c in dcn6_autoint4x32 five_nets; do
  timeout 400 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_$c.json 2> $O/bench_$c.err
done
DTB_DENSE_DIRECT=1 timeout 400 python bench.py --config dcn6_autoint4x32 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_dcn6_direct.json 2> $O/bench_dcn6_direct.err
tail -n 4 $O/pytest_sel.log $O/pytest_dense_direct.log; cat $O/pnn_once.log $O/dense_once_0.log $O/dense_once_1.log; for f in $O/bench_*.json; do cut -c1-200 $f; done
