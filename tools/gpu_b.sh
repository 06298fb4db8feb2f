#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
timeout 600 python tools/tune_adain.py 16 256 > gpurun_out/tune.log 2>&1
timeout 600 python tools/tune_adain.py 32 16 > gpurun_out/tune16.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log
tail -15 gpurun_out/pytest.log; cat gpurun_out/tune.log; tail -4 gpurun_out/tune16.log; cat gpurun_out/bench.log
