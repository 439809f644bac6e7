#!/bin/bash
# Round 2, run X (2 GPUs): the bench line (default layout cfg2) and one whole generation on the final build.
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r2x_n2.json 2> gpurun_out/bench_r2x_n2.err; grep '^{' gpurun_out/bench_r2x_n2.json | cut -c1-300; tail -2 gpurun_out/bench_r2x_n2.err
timeout 900 $TR bench.py --gpus 2 --workload video > gpurun_out/video_r2x_n2.json 2> gpurun_out/video_r2x_n2.err; grep '^{' gpurun_out/video_r2x_n2.json | cut -c1-1000; tail -2 gpurun_out/video_r2x_n2.err
