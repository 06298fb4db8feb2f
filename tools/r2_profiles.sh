#!/bin/bash
# Round-2 profile collection (GPU box): launch lists with tensor-pipe / DRAM metrics for every config, one full capture
# of the headline kernel and of the flash attention kernel.  Every ncu command runs only after the same command exited 0.
set -u
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
B="python bench.py --steps 2 --warmup 3 --no-extra --no-cpu --no-e2e --no-sustained"
$B > gpurun_out/p_bench_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv $B > gpurun_out/p_bench_ncu.log 2>&1
python tools/one_flash.py 128 1 > gpurun_out/p_flash_plain.log 2>&1 && ncu --metrics $M --clock-control none -c 100 --csv --log-file gpurun_out/r02_launches_flash.csv python tools/one_flash.py 128 1 > gpurun_out/p_flash_ncu.log 2>&1
python tools/one_wct.py 2 > gpurun_out/p_wct_plain.log 2>&1 && ncu --metrics $M --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_wct.csv python tools/one_wct.py 2 > gpurun_out/p_wct_ncu.log 2>&1
python tools/one_flash.py 128 1 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:flash_attn -s 2 -c 2 -o gpurun_out/r02_flash_full python tools/one_flash.py 128 1 > gpurun_out/p_flash_full.log 2>&1
$B > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:adain_tma -s 15 -c 5 -o gpurun_out/r02_adain_full $B > gpurun_out/p_adain_full.log 2>&1
python tools/one_wct.py 2 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:cov_tma -s 2 -c 1 -o gpurun_out/r02_cov_tma_full python tools/one_wct.py 2 > gpurun_out/p_cov_full.log 2>&1
python tools/one_wct.py 16 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ns_gemm -s 60 -c 2 -o gpurun_out/r02_ns_gemm_full python tools/one_wct.py 16 > gpurun_out/p_ns_full.log 2>&1
python tools/one_mrf.py > /dev/null 2>&1 && ncu --metrics $M --clock-control none -c 100 --csv --log-file gpurun_out/r02_launches_mrf.csv python tools/one_mrf.py > gpurun_out/p_mrf_ncu.log 2>&1
python tools/one_adaptive.py > /dev/null 2>&1 && ncu --metrics $M --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_adaptive.csv python tools/one_adaptive.py > gpurun_out/p_ada_ncu.log 2>&1
tail -2 gpurun_out/p_flash_full.log gpurun_out/p_adain_full.log gpurun_out/p_cov_full.log gpurun_out/p_ns_full.log
