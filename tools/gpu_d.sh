#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_adain_gpu.py -m gpu -q --timeout 60 > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
timeout 120 python tools/tune_adain.py 16 256 > gpurun_out/tune.log 2>&1; echo "tune exit $?" >> gpurun_out/tune.log
timeout 60 python tools/one_adain.py 8 256 1 4 > gpurun_out/one.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:adain_tma -s 2 -c 1 -o gpurun_out/adain_tma_blend python tools/one_adain.py 8 256 1 4 > gpurun_out/ncu1.log 2>&1
timeout 60 python tools/one_adain.py 8 256 0 4 > gpurun_out/one0.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:adain_tma -s 2 -c 1 -o gpurun_out/adain_tma_plain python tools/one_adain.py 8 256 0 4 > gpurun_out/ncu0.log 2>&1
tail -3 gpurun_out/pytest.log; cat gpurun_out/tune.log; tail -2 gpurun_out/ncu1.log
