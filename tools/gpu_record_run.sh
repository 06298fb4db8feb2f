#!/bin/bash
# round-1 record run: full GPU suite, bench (both arms), per-op bench, ncu launch list and full captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
tail -3 gpurun_out/pytest.log
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref exit $?"; tail -1 gpurun_out/bench_ref.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/bench.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench.log
timeout 900 python tools/bench_ops.py > gpurun_out/bench_ops.log 2>&1; cat gpurun_out/bench_ops.log | cut -c1-700
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launches.log 2>&1; echo "launch list exit $?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:seg_tma_kernel -s 2 -c 1 -o gpurun_out/seg_tma3 -f python tools/one_seg.py 256 > gpurun_out/ncu_seg.log 2>&1; echo "ncu seg exit $?"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:adain_tma_kernel --csv --log-file gpurun_out/bench_traffic.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_traffic.log 2>&1; echo "traffic exit $?"
timeout 300 python tools/one_adain.py 8 256 0 4 > /dev/null 2>&1 && timeout 600 ncu --set full --import-source on --clock-control none -k regex:adain_tma_kernel -s 2 -c 1 -o gpurun_out/adain_tma_plain -f python tools/one_adain.py 8 256 0 4 > gpurun_out/ncu_plain.log 2>&1; echo "ncu plain exit $?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:adain_tma_kernel -s 2 -c 1 -o gpurun_out/adain_tma_blend -f python tools/one_adain.py 8 256 1 4 > gpurun_out/ncu_blend.log 2>&1; echo "ncu blend exit $?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:pair_moments_kernel -c 1 -o gpurun_out/pair_moments -f python tools/bench_ops.py losses > gpurun_out/ncu_pair.log 2>&1; echo "ncu pair exit $?"
