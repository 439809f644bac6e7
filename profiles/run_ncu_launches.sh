#!/bin/bash
# Launch list of one denoising step (run under gpurun, 1 GPU): every launch of our kernels in the second step of
# `bench.py --quick`, device time only. Numbers printed under ncu are never bench values. Usage: run_ncu_launches.sh TAG
TAG=${1:-r1}
B="python bench.py --steps 1 --warmup 1 --quick --no-cpu-baseline"
K='regex:gemm|fmha_fwd_kernel|ln_kernel|rmsnorm_rope_kernel|patchify|unpatchify|linear_f32|sinusoid|im2col|groupnorm|nchw_to_nhwc'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 544 -c 560 --csv \
    --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/launches_$TAG.log 2>&1
tail -2 gpurun_out/launches_$TAG.log
