#!/bin/bash
mkdir -p gpurun_out
SEG_SHAPE=8,256,512,512 timeout 300 python tools/seg_sweep.py 16 24 32 48 64 128 256 > gpurun_out/seg_sweep512.log 2>&1; cat gpurun_out/seg_sweep512.log
