#!/bin/bash
mkdir -p gpurun_out
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size"
timeout 100 python tools/one_sanet.py 128 fp32 > gpurun_out/one_sanet.log 2>&1 &&
timeout 300 ncu --metrics $M --clock-control none -k regex:"gemm_packed|attn_rows|pack_operand" --csv --log-file gpurun_out/sanet_kernels.csv python tools/one_sanet.py 128 fp32 > gpurun_out/ncu_sanet.log 2>&1
timeout 100 python tools/one_wct.py 1 > gpurun_out/one_wct.log 2>&1 &&
timeout 300 ncu --metrics $M --clock-control none -k regex:"gemm_packed|pack_operand|jacobi" --csv --log-file gpurun_out/wct_kernels.csv python tools/one_wct.py 1 > gpurun_out/ncu_wct.log 2>&1
echo done
