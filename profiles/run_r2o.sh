#!/bin/bash
# Round 2, run O (8 GPUs): VAE decode in 8 bands, then one whole generation with the banded decode.
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571"
timeout 600 $TR bench.py --gpus 8 --workload vae --steps 2 > gpurun_out/vae_r2o_n8.json 2> gpurun_out/vae_r2o_n8.err; grep '^{' gpurun_out/vae_r2o_n8.json | cut -c1-900; tail -5 gpurun_out/vae_r2o_n8.err
timeout 900 $TR bench.py --gpus 8 --workload video > gpurun_out/video_r2o_n8.json 2> gpurun_out/video_r2o_n8.err; grep '^{' gpurun_out/video_r2o_n8.json | cut -c1-1200; tail -5 gpurun_out/video_r2o_n8.err
