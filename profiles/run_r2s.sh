#!/bin/bash
# Round 2, run S (1 GPU): validation of the final build: full GPU suite, smoke, the default bench line, the reference arm,
# one whole generation.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/pytest_r2s.log; cat gpurun_out/pytest_r2s.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 > gpurun_out/smoke_r2s.log; cat gpurun_out/smoke_r2s.log
timeout 900 python bench.py > gpurun_out/bench_r2s.json 2> gpurun_out/bench_r2s.err; grep '^{' gpurun_out/bench_r2s.json | cut -c1-400; tail -3 gpurun_out/bench_r2s.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r2s.json 2> gpurun_out/bench_ref_r2s.err; grep '^{' gpurun_out/bench_ref_r2s.json | cut -c1-400
timeout 900 python bench.py --workload video > gpurun_out/video_r2s_n1.json 2> gpurun_out/video_r2s_n1.err; grep '^{' gpurun_out/video_r2s_n1.json | cut -c1-900; tail -3 gpurun_out/video_r2s_n1.err
