#!/bin/bash
# round-2 call 20: whole GPU suite at HEAD (AFM / FiBiNet included) + per-kernel durations of the AFM / bilinear calls
O=gpurun_out/r2c20; mkdir -p $O
timeout 700 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
ROWS=16384 REPS=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file /tmp/f3_launches.csv python tools/f3_once.py > $O/ncu_f3.log 2>&1
python tools/summarize_launches.py /tmp/f3_launches.csv > $O/f3_launches.txt 2>&1
grep -E "passed|failed|FAILED|rc=" $O/pytest_gpu.log | head -20; cat $O/f3_launches.txt | head -20
