#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_adain_gpu.py tests/test_training_gpu.py tests/test_losses_gpu.py -m gpu -q --timeout 300 > gpurun_out/pytest_adain.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_adain.log
tail -4 gpurun_out/pytest_adain.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu > gpurun_out/bench_quick.log 2>&1; cat gpurun_out/bench_quick.log
