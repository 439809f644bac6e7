#!/bin/bash
# Round-1 (session 4) GPU recipe, one gpurun call, 1 GPU: parity tests, bench line, launch list, --set full captures.
# Numbers printed under ncu are never bench values.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2>/dev/null; cat gpurun_out/bench_ref.json
B="python bench.py --steps 1 --warmup 1 --quick --no-cpu-baseline"
K='regex:gemm|fmha_fwd_kernel|ln_kernel|rmsnorm_rope|patchify|unpatchify|linear_f32|sinusoid|im2col|groupnorm|nchw_to_nhwc|scatter|cast|swap'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 1300 --csv \
    --log-file gpurun_out/launches_r1b.csv $B > gpurun_out/launches_r1b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fmha_fwd_kernel -s 0 -c 2 \
    -o gpurun_out/prof_fmha_r1b -f $B > gpurun_out/prof_fmha_r1b.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:ln_kernel|rmsnorm_rope' -s 30 -c 5 \
    -o gpurun_out/prof_rows_r1b -f $B > gpurun_out/prof_rows_r1b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16_kernel -s 36 -c 6 \
    -o gpurun_out/prof_gemm_r1b -f $B > gpurun_out/prof_gemm_r1b.log 2>&1
ls -la gpurun_out/
