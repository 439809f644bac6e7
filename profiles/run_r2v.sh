#!/bin/bash
# Round 2, run V (1 GPU): distances TeaCache sees on the synthetic model, then a threshold that skips.
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --workload loop50 --teacache 0.5 > gpurun_out/loop50_tc_r2v_a.json 2> gpurun_out/loop50_tc_r2v_a.err; grep '^{' gpurun_out/loop50_tc_r2v_a.json | cut -c600-1300; tail -2 gpurun_out/loop50_tc_r2v_a.err
