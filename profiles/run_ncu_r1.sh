#!/bin/bash
# Round-1 profiling recipe (run under gpurun, 1 GPU). Numbers printed under ncu are never bench values.
set -x
B="python bench.py --steps 1 --warmup 1 --quick --no-cpu-baseline"
K='regex:gemm_bf16_kernel|fmha_fwd_kernel|ln_kernel|rmsnorm_rope_kernel|patchify|unpatchify|linear_f32|sinusoid|im2col|groupnorm|nchw_to_nhwc'
# every launch of our kernels in the second (timed) step, device time only
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 544 -c 560 --csv \
    --log-file gpurun_out/launches_r1.csv $B > gpurun_out/launches_r1.log 2>&1
# block-0 GEMMs (qkv, o, cross-q, cross-o, ffn.0, ffn.2): 46 earlier GEMM launches are front-end / static work
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 46 -c 6 \
    -o gpurun_out/prof_gemm_r1 -f $B > gpurun_out/prof_gemm_r1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fmha_fwd_kernel -s 0 -c 2 \
    -o gpurun_out/prof_fmha_r1 -f $B > gpurun_out/prof_fmha_r1.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:ln_kernel|rmsnorm_rope_kernel' -s 2 -c 6 \
    -o gpurun_out/prof_rows_r1 -f $B > gpurun_out/prof_rows_r1.log 2>&1
ls -la gpurun_out/
