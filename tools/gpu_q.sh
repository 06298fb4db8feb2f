#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
tail -40 gpurun_out/pytest.log
timeout 300 python tools/bench_ops.py losses sanet_bwd > gpurun_out/bench_losses.log 2>&1; cat gpurun_out/bench_losses.log
