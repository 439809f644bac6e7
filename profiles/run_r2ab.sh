#!/bin/bash
# Round 2, run AB (2 GPUs): the point-to-point form of the halo exchange under NCCL (FLEXAM_VAE_HALO=p2p), short clip.
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631"
FLEXAM_VAE_HALO=p2p timeout 240 $TR bench.py --gpus 2 --workload vae --vae-frames 3 --steps 1 > gpurun_out/vae_p2p_r2ab_n2.json 2> gpurun_out/vae_p2p_r2ab_n2.err; echo rc=$?; grep '^{' gpurun_out/vae_p2p_r2ab_n2.json | cut -c1-900; tail -3 gpurun_out/vae_p2p_r2ab_n2.err
