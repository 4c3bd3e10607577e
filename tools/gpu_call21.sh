#!/bin/bash
# round-2 call 21: AFM backward with every pair computed once (afm_rows_kernel<MODE 2> + afm_bwd_gather_kernel)
O=gpurun_out/r2c21; mkdir -p $O
timeout 300 python -m pytest tests/test_native_gpu.py tests/test_model_gpu.py tests/test_reference_golden.py -m gpu -q -k "afm" > $O/pytest_afm.log 2>&1; echo "rc=$?" >> $O/pytest_afm.log
DTB_AFM_BWD=1 timeout 300 python -m pytest tests/test_native_gpu.py -m gpu -q -k "afm" > $O/pytest_afm_mode1.log 2>&1; echo "rc=$?" >> $O/pytest_afm_mode1.log
timeout 200 python tools/f3_once.py > $O/f3_once.log 2>&1
grep -E "passed|failed|FAILED|rc=|Mismatch|Max " $O/pytest_afm.log $O/pytest_afm_mode1.log | head -30; grep afm $O/f3_once.log
