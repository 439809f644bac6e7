#!/bin/bash
# Round 2, run AE (1 GPU): last validation of the committed tree: full GPU suite, smoke, the default bench line.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/pytest_r2ae.log; cat gpurun_out/pytest_r2ae.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 > gpurun_out/smoke_r2ae.log; cat gpurun_out/smoke_r2ae.log
timeout 900 python bench.py > gpurun_out/bench_r2ae.json 2> gpurun_out/bench_r2ae.err; grep '^{' gpurun_out/bench_r2ae.json | cut -c1-250; tail -2 gpurun_out/bench_r2ae.err
