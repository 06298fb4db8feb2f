#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_adain_gpu.py tests/test_training_gpu.py tests/test_sanet_gpu.py -m gpu -q --timeout 300 > gpurun_out/pytest_adain.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_adain.log
tail -4 gpurun_out/pytest_adain.log
timeout 600 python tools/ab_twin.py > gpurun_out/ab_twin.log 2>&1; cat gpurun_out/ab_twin.log
