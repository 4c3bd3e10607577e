#!/bin/bash
# round-2 call 23: position-tiled FGCNN filter-gradient kernel (conv_fields_bwd_dw_tiled_kernel)
O=gpurun_out/r2c23; mkdir -p $O
export DTB_FGCNN_DW=1
timeout 200 python -m pytest tests/test_native_gpu.py tests/test_model_gpu.py -m gpu -q -k "fgcnn" > $O/pytest_fgcnn.log 2>&1; echo "rc=$?" >> $O/pytest_fgcnn.log
ONLY=fgcnn ROWS=65536 timeout 100 python tools/f3_once.py > $O/fgcnn_once.log 2>&1
grep -E "passed|failed|FAILED|rc=|Mismatch|Max " $O/pytest_fgcnn.log | head; cat $O/fgcnn_once.log
