#!/bin/bash
# first GPU contact: tests, smoke, tuning sweep, bench, ncu launch list + one full capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python tools/tune_adain.py 16 256 > gpurun_out/tune.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log
timeout 300 python bench.py --steps 2 --warmup 3 --batch 8 --no-e2e --no-cpu > gpurun_out/bench_small.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --batch 8 --no-e2e --no-cpu > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:adain_pipe -s 15 -c 2 -o gpurun_out/adain_pipe \
    python bench.py --steps 2 --warmup 3 --batch 8 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/pytest.log; cat gpurun_out/smoke.log; tail -3 gpurun_out/tune.log; cat gpurun_out/bench.log
