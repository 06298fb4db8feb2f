#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_adain_gpu.py -m gpu -q --timeout 60 > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
timeout 120 python tools/tune_adain.py 16 256 > gpurun_out/tune.log 2>&1; echo "tune exit $?" >> gpurun_out/tune.log
tail -5 gpurun_out/pytest.log; cat gpurun_out/tune.log
