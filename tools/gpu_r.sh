#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_seg_gpu.py -m gpu -q --timeout 300 > gpurun_out/pytest_seg.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_seg.log
tail -5 gpurun_out/pytest_seg.log
timeout 300 python tools/seg_sweep.py 128 192 256 384 512 1024 > gpurun_out/seg_sweep.log 2>&1; cat gpurun_out/seg_sweep.log
