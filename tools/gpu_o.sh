#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_mrf_gpu.py tests/test_wct_gpu.py tests/test_sanet_gpu.py -m gpu -q --timeout 120 > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
timeout 400 python tools/bench_ops.py wct sanet mrf > gpurun_out/bench_ops.log 2>&1
tail -3 gpurun_out/pytest.log; cat gpurun_out/bench_ops.log
bash tools/gpu_n.sh
