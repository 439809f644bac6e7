#!/bin/bash
# Round-1 (session 5) GPU recipe, one gpurun call, 1 GPU: parity tests, bench line (both arms), launch list of the
# final round-1 build, --set full captures of block-0 GEMMs / attention / row kernels. Numbers printed under ncu are
# never bench values.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_e.log 2>&1; tail -5 gpurun_out/pytest_gpu_e.log
timeout 600 python bench.py > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; tail -c 3000 gpurun_out/bench_e.json
timeout 300 python tests/gpu_microbench.py gemm rows fmha > gpurun_out/micro_e.log 2>&1; cat gpurun_out/micro_e.log
B="python bench.py --steps 1 --warmup 1 --quick --no-cpu-baseline"
K='regex:gemm|fmha_fwd_kernel|ln_kernel|rmsnorm_rope|patchify|unpatchify|linear_f32|sinusoid|im2col|groupnorm|nchw_to_nhwc|scatter|cast|swap'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 1300 --csv \
    --log-file gpurun_out/launches_r1e.csv $B > gpurun_out/launches_r1e.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fmha_fwd_kernel -s 0 -c 2 \
    -o gpurun_out/prof_fmha_r1e -f $B > gpurun_out/prof_fmha_r1e.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:ln_kernel|rmsnorm_rope' -s 30 -c 5 \
    -o gpurun_out/prof_rows_r1e -f $B > gpurun_out/prof_rows_r1e.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16_kernel -s 36 -c 6 \
    -o gpurun_out/prof_gemm_r1e -f $B > gpurun_out/prof_gemm_r1e.log 2>&1
ls -la gpurun_out/
