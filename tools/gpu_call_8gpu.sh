#!/bin/bash
# round-2 8-GPU call: NCCL replica tests at world 8 (world 2: profiles/r2_dp_nccl_tests_w2.log), bench at N = 8 and
# N = 2, kernel timeline of one 8-GPU step
O=gpurun_out/r2_8gpu; mkdir -p $O
timeout 600 python -m pytest tests/test_dp_gpu.py -m gpu -q -s -k "8]" > $O/dp_nccl_w8.log 2>&1; echo "rc=$?" >> $O/dp_nccl_w8.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_n8.json 2> $O/bench_n8.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 tools/dp_timeline.py > $O/dp_timeline.log 2>&1
cp gpurun_out/dp_timeline_w8.txt $O/ 2>/dev/null
tail -n 8 $O/dp_nccl_w8.log; for n in 2 8; do cut -c1-260 $O/bench_n$n.json; done; tail -n 3 $O/bench_n8.err $O/dp_timeline.log; head -n 3 $O/dp_timeline_w8.txt
