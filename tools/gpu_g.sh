#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_sanet_gpu.py -m gpu -q --timeout 90 > gpurun_out/pytest_sanet.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_sanet.log
tail -40 gpurun_out/pytest_sanet.log
