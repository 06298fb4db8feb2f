#!/bin/bash
# full round-1 measurement: all GPU tests, smoke, bench (both arms), ncu launch list + full captures
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 200 python tools/tune_adain.py 16 256 > gpurun_out/tune.log 2>&1; echo "tune exit $?" >> gpurun_out/tune.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_ref.log
timeout 120 python bench.py --steps 2 --warmup 3 --batch 8 --no-e2e --no-cpu > gpurun_out/bench_small.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --batch 8 --no-e2e --no-cpu > gpurun_out/ncu_launches.log 2>&1
timeout 60 python tools/one_adain.py 8 256 1 4 > gpurun_out/one.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:adain_tma -s 2 -c 1 -o gpurun_out/adain_tma_blend python tools/one_adain.py 8 256 1 4 > gpurun_out/ncu1.log 2>&1
timeout 60 python tools/one_adain.py 8 256 0 4 > gpurun_out/one0.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:adain_tma -s 2 -c 1 -o gpurun_out/adain_tma_plain python tools/one_adain.py 8 256 0 4 > gpurun_out/ncu0.log 2>&1
tail -3 gpurun_out/pytest.log; cat gpurun_out/smoke.log; tail -4 gpurun_out/tune.log; cat gpurun_out/bench.log; cat gpurun_out/bench_ref.log
