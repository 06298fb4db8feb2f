#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_seg_gpu.py -m gpu -q --timeout 90 > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
timeout 200 python tools/bench_ops.py seg > gpurun_out/bench_ops2.log 2>&1
timeout 100 python tools/one_seg.py > gpurun_out/one_seg.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:seg_tma -s 1 -c 1 -o gpurun_out/seg_tma python tools/one_seg.py > gpurun_out/ncu_seg.log 2>&1
tail -3 gpurun_out/pytest.log; cat gpurun_out/bench_ops2.log
