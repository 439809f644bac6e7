#!/bin/bash
# Round 2, run Y (1 GPU): tensor-map descriptor cache: full GPU suite, then the bench line (host enqueue time per step).
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/pytest_r2y.log; cat gpurun_out/pytest_r2y.log
timeout 900 python bench.py --no-library-baseline --no-cpu-baseline > gpurun_out/bench_r2y.json 2> gpurun_out/bench_r2y.err; grep '^{' gpurun_out/bench_r2y.json | cut -c1-300; tail -3 gpurun_out/bench_r2y.err
