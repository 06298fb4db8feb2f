#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_adain_gpu.py tests/test_training_gpu.py tests/test_seg_gpu.py -m gpu -q --timeout 300 > gpurun_out/pytest_adain.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_adain.log
tail -3 gpurun_out/pytest_adain.log
timeout 300 python tools/bench_ops.py train5
timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu > gpurun_out/bench_quick.log 2>&1; grep -o '"value": [0-9.]*, "unit": "images/s", "n_gpus"\|"achieved": [0-9.]*\|"per_level_GBs": {[^}]*}\|"sm_mhz": [0-9.]*' gpurun_out/bench_quick.log
