#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
tail -15 gpurun_out/pytest.log
timeout 300 python tools/bench_ops.py losses > gpurun_out/bench_losses.log 2>&1; cat gpurun_out/bench_losses.log
