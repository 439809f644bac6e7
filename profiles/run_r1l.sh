#!/bin/bash
# Round-1 (session 8) final GPU recipe, one gpurun call, 1 GPU: in-step comparison of the attention pipelines, the whole
# parity suite (no -x), smoke(), the bench line, kernel microbenchmarks, launch list and --set full captures of the
# default build. Numbers printed under ncu are never bench values.
set -x
mkdir -p gpurun_out
Q="python bench.py --steps 3 --warmup 3 --quick --no-cpu-baseline"
: > gpurun_out/instep_pipes_r1l.log
for cfg in "3 0 0" "2 0 0" "1 1 2"; do
  set -- $cfg
  echo "== in-step FX_FMHA_PIPE=$1 FX_FMHA_TOKEN=$2 FX_FMHA_POLY=$3" >> gpurun_out/instep_pipes_r1l.log
  FX_FMHA_PIPE=$1 FX_FMHA_TOKEN=$2 FX_FMHA_POLY=$3 timeout 300 $Q >> gpurun_out/instep_pipes_r1l.log 2>&1
done
cat gpurun_out/instep_pipes_r1l.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_l.log 2>&1; tail -15 gpurun_out/pytest_gpu_l.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_l.log 2>&1; tail -2 gpurun_out/smoke_l.log
timeout 600 python bench.py > gpurun_out/bench_l.json 2> gpurun_out/bench_l.err; tail -c 3500 gpurun_out/bench_l.json
timeout 300 python tests/gpu_microbench.py fmha gemm rows > gpurun_out/micro_l.log 2>&1; cat gpurun_out/micro_l.log
B="python bench.py --steps 1 --warmup 1 --quick --no-cpu-baseline"
K='regex:gemm|fmha|ln_kernel|rmsnorm_rope|patchify|unpatchify|linear_f32|sinusoid|im2col|groupnorm|nchw_to_nhwc|scatter|cast|swap'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 1300 --csv \
    --log-file gpurun_out/launches_r1l.csv $B > gpurun_out/launches_r1l.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fmha -s 0 -c 2 \
    -o gpurun_out/prof_fmha_r1l -f $B > gpurun_out/prof_fmha_r1l.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16_kernel -s 36 -c 6 \
    -o gpurun_out/prof_gemm_r1l -f $B > gpurun_out/prof_gemm_r1l.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:ln_kernel|rmsnorm_rope' -s 30 -c 5 \
    -o gpurun_out/prof_rows_r1l -f $B > gpurun_out/prof_rows_r1l.log 2>&1
ls -la gpurun_out/ | tail -20
