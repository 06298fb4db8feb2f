#!/bin/bash
mkdir -p gpurun_out
timeout 100 python tools/one_wct.py 2 > gpurun_out/one_wct.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/wct_launches.csv python tools/one_wct.py 2 > gpurun_out/ncu_wct.log 2>&1
timeout 100 python tools/one_seg.py > gpurun_out/one_seg.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:seg_pipe -s 1 -c 1 -o gpurun_out/seg_pipe python tools/one_seg.py > gpurun_out/ncu_seg.log 2>&1
tail -2 gpurun_out/ncu_wct.log
