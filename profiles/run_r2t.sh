#!/bin/bash
# Round 2, run T (8 GPUs): one whole generation on the final build (banded decode, 160-wide encoder tiles).
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581"
timeout 900 $TR bench.py --gpus 8 --workload video > gpurun_out/video_r2t_n8.json 2> gpurun_out/video_r2t_n8.err; grep '^{' gpurun_out/video_r2t_n8.json | cut -c1-1200; tail -3 gpurun_out/video_r2t_n8.err
timeout 600 $TR bench.py --gpus 8 --workload vae --steps 2 > gpurun_out/vae_r2t_n8.json 2> gpurun_out/vae_r2t_n8.err; grep '^{' gpurun_out/vae_r2t_n8.json | cut -c1-900
