#!/bin/bash
O=gpurun_out/r2c9; mkdir -p $O
timeout 200 python tools/debug_five.py > $O/debug_five_p0.log 2>&1
timeout 200 python tools/debug_wgrad.py > $O/debug_wgrad.log 2>&1
PREC=4 CHECKB=1 REPS=3 timeout 300 python tools/cin_once.py > $O/cin_once_p4.log 2>&1
grep -vE "UserWarning|Consider|print" $O/debug_five_p0.log | head -n 30; cat $O/debug_wgrad.log; tail -n 3 $O/cin_once_p4.log
