#!/bin/bash
# Round 2, run J (1 GPU): what the driver runs at round end — the whole GPU suite, smoke(), the default bench line, the
# reference arm — on the final build.
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r2j.log 2>&1; tail -6 gpurun_out/pytest_gpu_r2j.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2j.log 2>&1; tail -2 gpurun_out/smoke_r2j.log
timeout 900 python bench.py > gpurun_out/bench_r2j.json 2> gpurun_out/bench_r2j.err; tail -c 1500 gpurun_out/bench_r2j.json; tail -3 gpurun_out/bench_r2j.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r2j.json 2> gpurun_out/bench_ref_r2j.err; tail -c 700 gpurun_out/bench_ref_r2j.json
