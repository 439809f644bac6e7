#!/bin/bash
# Round 2, run U (2 GPUs): the 50-step loop with TeaCache and cfg_skip enabled, one GPU vs two (CFG-branch parallel and
# sequence parallel): the residual caches live on different ranks there; final-latent checksums must match.
set -x
mkdir -p gpurun_out
A="--workload loop50 --teacache ${TC:-0.1} --cfg-skip 0.25"
timeout 600 python bench.py $A > gpurun_out/loop50_tc${TC:-}_r2u_n1.json 2> gpurun_out/loop50_tc${TC:-}_r2u_n1.err; grep '^{' gpurun_out/loop50_tc${TC:-}_r2u_n1.json | cut -c1-1100; tail -2 gpurun_out/loop50_tc${TC:-}_r2u_n1.err
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29591"
timeout 600 $TR bench.py --gpus 2 $A > gpurun_out/loop50_tc${TC:-}_r2u_cfg2.json 2> gpurun_out/loop50_tc${TC:-}_r2u_cfg2.err; grep '^{' gpurun_out/loop50_tc${TC:-}_r2u_cfg2.json | cut -c1-1100; tail -2 gpurun_out/loop50_tc${TC:-}_r2u_cfg2.err
FLEXAM_CFG_SIZE=1 timeout 600 $TR bench.py --gpus 2 $A > gpurun_out/loop50_tc${TC:-}_r2u_sp2.json 2> gpurun_out/loop50_tc${TC:-}_r2u_sp2.err; grep '^{' gpurun_out/loop50_tc${TC:-}_r2u_sp2.json | cut -c1-1100; tail -2 gpurun_out/loop50_tc${TC:-}_r2u_sp2.err
