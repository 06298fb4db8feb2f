#!/bin/bash
# copy the outputs of tools/r2_profiles.sh + the final bench / timing logs from gpurun_out/ into profiles/ and regenerate
# the summaries (run in the dev container after the GPU call)
set -u
cd "$(dirname "$0")/.."
cp gpurun_out/r02_launches_bench.csv gpurun_out/r02_launches_flash.csv gpurun_out/r02_launches_wct.csv gpurun_out/r02_launches_mrf.csv gpurun_out/r02_launches_adaptive.csv profiles/
cp gpurun_out/r02_bench_n1_final.json profiles/r02_bench_n1.json
cp gpurun_out/r02_bench_ref_final.json profiles/r02_bench_reference_arm.json
cp gpurun_out/r02_bench_ops_final.jsonl profiles/r02_bench_ops.jsonl
cp gpurun_out/r02_wct_time_final.log profiles/r02_wct_time.log
cp gpurun_out/r02_cov_prof.log gpurun_out/r02_roots_ab.jsonl gpurun_out/r02_pw_ab.jsonl profiles/ 2>/dev/null
for k in cov_tma ns_gemm pwconv; do python tools/ncu_summary.py gpurun_out/r02_${k}_full.ncu-rep > profiles/r02_${k}_ncu_summary.txt 2>&1; done
python tools/ncu_summary.py gpurun_out/r02_flash_full.ncu-rep > profiles/r02_flash_attn_ncu_summary.txt 2>&1
python tools/ncu_summary.py gpurun_out/r02_adain_full.ncu-rep > profiles/r02_adain_tma_ncu_summary.txt 2>&1
python tools/r2_tensor_pipe.py > /dev/null
