#!/bin/bash
# Round-1 profiling recipe (run under gpurun, 1 GPU). Numbers printed under ncu are never bench values.
#   launches_r1.csv      every launch of our kernels in two denoising steps of `bench.py --quick` (the second step is
#                        the one summarised), --metrics gpu__time_duration.sum --clock-control none
#   prof_*_r1.ncu-rep    --set full captures of block 0's GEMMs, the two attention launches and the row kernels
set -x
B="python bench.py --steps 1 --warmup 1 --quick --no-cpu-baseline"
K='regex:gemm|fmha_fwd_kernel|ln_kernel|rmsnorm_rope_kernel|patchify|unpatchify|linear_f32|sinusoid|im2col|groupnorm|nchw_to_nhwc'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 1100 --csv \
    --log-file gpurun_out/launches_r1.csv $B > gpurun_out/launches_r1.log 2>&1
# block-0 GEMMs of the first step (qkv, self o, cross q, cross o, ffn.0, ffn.2): the 36 earlier CTA-pair GEMM launches are the
# CNN fuser, text embedding, cross K/V and patch/ref embedding
ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16_kernel -s 36 -c 6 \
    -o gpurun_out/prof_gemm_r1 -f $B > gpurun_out/prof_gemm_r1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fmha_fwd_kernel -s 0 -c 2 \
    -o gpurun_out/prof_fmha_r1 -f $B > gpurun_out/prof_fmha_r1.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:ln_kernel|rmsnorm_rope_kernel' -s 30 -c 5 \
    -o gpurun_out/prof_rows_r1 -f $B > gpurun_out/prof_rows_r1.log 2>&1
ls -la gpurun_out/
