#!/bin/bash
# Round 2, run W (8 GPUs): the 50-step loop with TeaCache + cfg_skip in the cfg2 x sp4 layout (checksum must equal N = 1, 2).
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29601"
timeout 600 $TR bench.py --gpus 8 --workload loop50 --teacache 2.0 --cfg-skip 0.25 > gpurun_out/loop50_tc2.0_r2w_n8.json 2> gpurun_out/loop50_tc2.0_r2w_n8.err; grep '^{' gpurun_out/loop50_tc2.0_r2w_n8.json | cut -c1-1300; tail -2 gpurun_out/loop50_tc2.0_r2w_n8.err
