#!/bin/bash
O=gpurun_out/r2c13; mkdir -p $O
DTB_DEBUG_CAPTURE=1 timeout 300 python tools/debug_capture2.py > $O/debug_capture2.log 2>&1
cat $O/debug_capture2.log | cut -c1-400
