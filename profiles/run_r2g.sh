#!/bin/bash
# Round 2, run G (2 GPUs): implicit-GEMM convolution in the control fuser — operator tests, whole-step parity, bench
# line at N=1 and in the SP layout (checksums must agree), and the ncu launch list of the round-2 step.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s -k "conv or golden or benchmarked or bf16_policy or deterministic" > gpurun_out/pytest_gpu_r2g.log 2>&1; tail -25 gpurun_out/pytest_gpu_r2g.log
timeout 600 python bench.py --steps 5 --warmup 3 --itemise --no-library-baseline --no-cpu-baseline > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err; tail -c 900 gpurun_out/bench_r2g.json; tail -3 gpurun_out/bench_r2g.err
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531"
FLEXAM_CFG_SIZE=1 timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r2g_sp2.json 2> gpurun_out/bench_r2g_sp2.err; tail -c 900 gpurun_out/bench_r2g_sp2.json; tail -3 gpurun_out/bench_r2g_sp2.err
B="python bench.py --steps 1 --warmup 1 --quick --no-cpu-baseline"
K='regex:gemm|fmha|ln_kernel|rmsnorm_rope|patchify|unpatchify|linear_f32|sinusoid|im2col|groupnorm|nchw_to_nhwc|scatter|cast|swap|dedup|fingerprint|modulation|split_planes'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 1300 --csv \
    --log-file gpurun_out/launches_r2g.csv $B > gpurun_out/launches_r2g.log 2>&1
tail -2 gpurun_out/launches_r2g.log
