#!/bin/bash
# Round 2, run L (8 GPUs): one whole generation on the box (encodes sharded over ranks, loop cfg2 x sp4, decode on rank 0)
# and the bench line of the final build.
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $TR bench.py --gpus 8 --workload video > gpurun_out/video_r2l_n8.json 2> gpurun_out/video_r2l_n8.err; grep '^{' gpurun_out/video_r2l_n8.json; tail -3 gpurun_out/video_r2l_n8.err
timeout 600 $TR bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_r2l_n8.json 2> gpurun_out/bench_r2l_n8.err; grep '^{' gpurun_out/bench_r2l_n8.json | cut -c1-1200; tail -3 gpurun_out/bench_r2l_n8.err
