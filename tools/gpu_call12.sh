#!/bin/bash
O=gpurun_out/r2c12; mkdir -p $O
timeout 300 python tools/debug_capture2.py > $O/debug_capture2.log 2>&1
timeout 600 python -m pytest tests/test_native_gpu.py -m gpu -q -rfE -k "long_gap or full_batch or cin_fwd_bwd" > $O/pytest_sel.log 2>&1
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -rfE -k "pinned_host" > $O/pytest_new.log 2>&1
cat $O/debug_capture2.log | cut -c1-300; grep -E "passed|failed|^E  |FAILED" $O/pytest_sel.log | cut -c1-500 | head -20; grep -E "passed|failed|^E  |FAILED" $O/pytest_new.log | cut -c1-300 | head
