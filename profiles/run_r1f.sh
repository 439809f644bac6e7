#!/bin/bash
# Round-1 (session 8) GPU recipe, one gpurun call, 1 GPU: the shared-score-buffer attention pipeline (FX_FMHA_PIPE=2|3)
# against the in-place one (1): native timing + accuracy + whole-output checksum at full, ragged and tiny lengths,
# timelines, then the parity suite and the bench line with the fastest correct pipeline, launch list and --set full
# capture. Numbers printed under ncu are never bench values.
set -x
mkdir -p gpurun_out
N=tests/native
L=gpurun_out/fmha_pipes_r1f.log
python profiles/r1_tools/fmha_pipe_sweep.py $L > gpurun_out/chosen_r1f.sh
cat gpurun_out/chosen_r1f.sh
source gpurun_out/chosen_r1f.sh
for pipe in 2 3; do
  echo "== fmha_trace pipe=$pipe" >> $L
  FX_FMHA_PIPE=$pipe timeout 120 $N/fmha_trace >> $L 2>&1
done
grep -A30 "== fmha_trace" $L | cut -c1-150
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_f.log 2>&1; tail -5 gpurun_out/pytest_gpu_f.log
timeout 600 python bench.py > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err; tail -c 3000 gpurun_out/bench_f.json
timeout 200 python tests/gpu_microbench.py fmha > gpurun_out/micro_f.log 2>&1; cat gpurun_out/micro_f.log
B="python bench.py --steps 1 --warmup 1 --quick --no-cpu-baseline"
K='regex:gemm|fmha|ln_kernel|rmsnorm_rope|patchify|unpatchify|linear_f32|sinusoid|im2col|groupnorm|nchw_to_nhwc|scatter|cast|swap'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 1300 --csv \
    --log-file gpurun_out/launches_r1f.csv $B > gpurun_out/launches_r1f.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fmha -s 0 -c 2 \
    -o gpurun_out/prof_fmha_r1f -f $B > gpurun_out/prof_fmha_r1f.log 2>&1
ls -la gpurun_out/
