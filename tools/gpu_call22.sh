#!/bin/bash
# round-2 call 22: FGCNN (convolution / pooling along the fields, tanh Dense) kernels, layer and nets; focal losses
O=gpurun_out/r2c22; mkdir -p $O
export DTB_TEST_FGCNN=1
timeout 400 python -m pytest tests/test_native_gpu.py -m gpu -q -k "fgcnn or tanh or focal" > $O/pytest_kernels.log 2>&1; echo "rc=$?" >> $O/pytest_kernels.log
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_reference_golden.py -m gpu -q -k "fg or focal" > $O/pytest_models.log 2>&1; echo "rc=$?" >> $O/pytest_models.log
grep -E "passed|failed|FAILED|rc=|Error|Mismatch|Max |err_msg|^E  " $O/pytest_kernels.log | head -40; grep -E "passed|failed|FAILED|rc=|Error:|Mismatch|Max " $O/pytest_models.log | head -40
ONLY=fgcnn timeout 200 python tools/f3_once.py > $O/fgcnn_once.log 2>&1; tail -4 $O/fgcnn_once.log
unset DTB_TEST_FGCNN
timeout 700 python -m pytest tests -m gpu -q > $O/pytest_gpu_all.log 2>&1; echo "rc=$?" >> $O/pytest_gpu_all.log; tail -n 3 $O/pytest_gpu_all.log
