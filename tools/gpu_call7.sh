#!/bin/bash
# round-2 GPU call 7 (2 GPUs): full suite with precision auto = fp16 single pass (incl. the NCCL replica tests at world 2),
# CIN timings, bench at N = 1 and N = 2
O=gpurun_out/r2c7; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -rfEs --timeout 900 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 600 python -m pytest tests/test_dp_gpu.py -m gpu -q -s > $O/dp_nccl_w2.log 2>&1; echo "rc=$?" >> $O/dp_nccl_w2.log
PREC=4 CHECKB=1 REPS=3 timeout 300 python tools/cin_once.py > $O/cin_once_p4.log 2>&1
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err
grep -E "passed|failed" $O/pytest.log | tail -n 2; grep -E "^FAILED|^ERROR" $O/pytest.log | head -n 30
tail -n 6 $O/dp_nccl_w2.log; tail -n 3 $O/cin_once_p4.log; cut -c1-250 $O/bench_n1.json; cut -c1-250 $O/bench_n2.json; tail -n 3 $O/bench_n2.err
