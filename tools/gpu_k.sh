#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_seg_gpu.py tests/test_wct_gpu.py -m gpu -q --timeout 90 > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
timeout 300 python tools/bench_ops.py seg wct > gpurun_out/bench_ops2.log 2>&1; echo "ops exit $?" >> gpurun_out/bench_ops2.log
tail -25 gpurun_out/pytest.log; cat gpurun_out/bench_ops2.log
