#!/bin/bash
O=gpurun_out/r2c11; mkdir -p $O
DTB_DEBUG_CAPTURE=1 timeout 300 python tools/debug_capture.py > $O/debug_capture.log 2>&1
timeout 600 python -m pytest tests/test_native_gpu.py -m gpu -q -rfE -k "long_gap or full_batch or cin_fwd_bwd" > $O/pytest_sel.log 2>&1
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -rfE -k "cross_validation or pinned_host" > $O/pytest_new.log 2>&1
cat $O/debug_capture.log | cut -c1-700; grep -E "passed|failed|^E  |FAILED" $O/pytest_sel.log | cut -c1-400 | head -20; grep -E "passed|failed|^E  |FAILED" $O/pytest_new.log | cut -c1-300 | head
