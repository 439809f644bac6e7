#!/bin/bash
# Round 2, run E (1 GPU): parity suite incl. the umT5 encoder, text-encoder bench leg, attention microbench next to the
# library SDPA (burst and sustained), ncu launch list of the round-2 step and --set full captures of the top kernels.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu_r2e.log 2>&1; tail -30 gpurun_out/pytest_gpu_r2e.log
timeout 600 python bench.py --workload t5 --steps 10 --warmup 3 > gpurun_out/t5_r2e.json 2> gpurun_out/t5_r2e.err; cat gpurun_out/t5_r2e.json; tail -3 gpurun_out/t5_r2e.err
timeout 400 python tests/gpu_microbench.py fmha gemm > gpurun_out/micro_r2e.log 2>&1; cat gpurun_out/micro_r2e.log
B="python bench.py --steps 1 --warmup 1 --quick --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv \
    --log-file gpurun_out/launches_r2e.csv $B > gpurun_out/launches_r2e.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fmha -s 0 -c 2 \
    -o gpurun_out/prof_fmha_r2e -f $B > gpurun_out/prof_fmha_r2e.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16_kernel -s 36 -c 6 \
    -o gpurun_out/prof_gemm_r2e -f $B > gpurun_out/prof_gemm_r2e.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:ln_kernel|rmsnorm_rope' -s 30 -c 5 \
    -o gpurun_out/prof_rows_r2e -f $B > gpurun_out/prof_rows_r2e.log 2>&1
ls -la gpurun_out | tail -12
