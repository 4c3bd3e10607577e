#!/bin/bash
# one-shot GPU validation of the compact saved-activation format (run through gpurun)
mkdir -p gpurun_out
timeout -s KILL 110 python -m pytest tests/test_native_gpu.py tests/test_model_gpu.py -x -q --tb=short -k "tensor_core_backward or forward_and_training or final_before" > gpurun_out/compact_tests.log 2>&1
tail -4 gpurun_out/compact_tests.log
timeout -s KILL 60 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_compact.json 2> gpurun_out/bench_compact.err
cut -c1-330 gpurun_out/bench_compact.json
