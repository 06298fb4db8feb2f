#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/seg_sweep.py 96 256 > gpurun_out/seg_sweep.log 2>&1; cat gpurun_out/seg_sweep.log
timeout 300 python tools/one_seg.py 128 > gpurun_out/one_seg.log 2>&1 && \
timeout 600 ncu --set full --import-source on --clock-control none -k regex:seg_tma_kernel -s 2 -c 1 -o gpurun_out/seg_tma2 -f python tools/one_seg.py 128 > gpurun_out/ncu_seg.log 2>&1
tail -3 gpurun_out/ncu_seg.log
