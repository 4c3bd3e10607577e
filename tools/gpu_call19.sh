#!/bin/bash
# round-2 call 19: AFM and FiBiNet (SENET + BilinearInteraction) kernels, layers and nets
O=gpurun_out/r2c19; mkdir -p $O
timeout 600 python -m pytest tests/test_native_gpu.py -m gpu -q -k "afm or bilinear or senet" > $O/pytest_kernels.log 2>&1; echo "rc=$?" >> $O/pytest_kernels.log
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_reference_golden.py -m gpu -q -k "afm or fibi" > $O/pytest_models.log 2>&1; echo "rc=$?" >> $O/pytest_models.log
timeout 300 python tools/f3_once.py > $O/f3_once.log 2>&1
grep -E "passed|failed|Error|error|assert|FAILED|rc=" $O/pytest_kernels.log | head -40; grep -E "passed|failed|Error|FAILED|rc=|Mismatch|Max " $O/pytest_models.log | head -40; cat $O/f3_once.log | tail -12
