#!/bin/bash
# round-2 GPU call 2: tcgen05 Dense GEMMs (no cuBLAS in the .so), wide Cross, bounded Adam replay; the suite again with the
# CIN forward forced to the fp16 single pass; bench with a fresh batch per step
O=gpurun_out/r2c2; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -rfE --timeout 600 -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
DTB_CIN_PRECISION=4 timeout 900 python -m pytest tests -m gpu -q -rfE --timeout 600 > $O/pytest_p4.log 2>&1; echo "pytest rc=$?" >> $O/pytest_p4.log
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_default.json 2> $O/bench_default.err
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --cin-precision 4 > $O/bench_p4.json 2> $O/bench_p4.err
tail -n 8 $O/pytest.log; tail -n 25 $O/pytest_p4.log; cat $O/bench_*.json | cut -c1-300; tail -n 5 $O/*.err
