#!/bin/bash
# GPU call 2 (1 GPU): attention softmax A/B (lagging vs synchronous row maximum, exp2 split), pipe-rate tools,
# parity tests, quick bench. Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_c.log 2>&1; tail -5 gpurun_out/pytest_gpu_c.log
( for m in sync lag; do for p in 2 3 4; do FX_FMHA_MAX=$m FX_FMHA_POLY=$p python tests/gpu_microbench.py fmha | sed "s/^/$m /"; done; done ) > gpurun_out/fmha_ab.log 2>&1
cat gpurun_out/fmha_ab.log
python tests/gpu_microbench.py gemm rows > gpurun_out/micro_c.log 2>&1; cat gpurun_out/micro_c.log
./tests/native/fmha_trace > gpurun_out/fmha_trace_lag.log 2>&1; tail -20 gpurun_out/fmha_trace_lag.log
FX_FMHA_MAX=sync ./tests/native/fmha_trace > gpurun_out/fmha_trace_sync.log 2>&1
./tests/native/pipe_rate > gpurun_out/pipe_rate.log 2>&1; tail -40 gpurun_out/pipe_rate.log
./tests/native/umma_rate > gpurun_out/umma_rate.log 2>&1; tail -30 gpurun_out/umma_rate.log
python bench.py --no-cpu-baseline > gpurun_out/bench_c.json 2>gpurun_out/bench_c.err; cat gpurun_out/bench_c.json
