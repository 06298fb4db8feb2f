#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_mrf_gpu.py -m gpu -q --timeout 60 -x > gpurun_out/pytest_mrf.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_mrf.log
tail -30 gpurun_out/pytest_mrf.log
