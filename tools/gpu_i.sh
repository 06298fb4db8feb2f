#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mrf_gpu.py tests/test_wct_gpu.py tests/test_sanet_gpu.py -m gpu -q --timeout 120 > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
timeout 600 python tools/bench_ops.py wct sanet mrf > gpurun_out/bench_ops.log 2>&1; echo "ops exit $?" >> gpurun_out/bench_ops.log
timeout 100 python tools/bench_ops.py mrf > gpurun_out/mrf_plain.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_packed -s 2 -c 1 -o gpurun_out/gemm_packed_mrf python tools/bench_ops.py mrf > gpurun_out/ncu_gemm.log 2>&1
tail -3 gpurun_out/pytest.log; cat gpurun_out/bench_ops.log
