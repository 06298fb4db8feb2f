#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
timeout 600 python tools/bench_ops.py > gpurun_out/bench_ops.log 2>&1; echo "ops exit $?" >> gpurun_out/bench_ops.log
# DRAM traffic of the actual bench launches (batch 32): metrics-only pass, 5 launches of one step
timeout 200 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_1step.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:adain_tma -s 15 -c 5 --csv --log-file gpurun_out/bench_traffic.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_traffic.log 2>&1
# tensor-core kernel capture (MRF NCC GEMM, bf16x3)
timeout 100 python tools/bench_ops.py mrf > gpurun_out/mrf_plain.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_packed -s 2 -c 1 -o gpurun_out/gemm_packed_mrf python tools/bench_ops.py mrf > gpurun_out/ncu_gemm.log 2>&1
tail -3 gpurun_out/pytest.log; cat gpurun_out/bench_ops.log; tail -8 gpurun_out/bench_traffic.csv | cut -c1-250
