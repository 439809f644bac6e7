#!/bin/bash
# Round 2, run K (1 GPU): one whole generation (umT5 + 8 VAE encodes + 50-step loop + VAE decode), every model native.
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --workload video > gpurun_out/video_r2k_n1.json 2> gpurun_out/video_r2k_n1.err; cat gpurun_out/video_r2k_n1.json; tail -5 gpurun_out/video_r2k_n1.err
