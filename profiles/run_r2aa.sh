#!/bin/bash
# Round 2, run AA (4 GPUs): multi-GPU parity check of the step (tests/gpu_dist_check.py) on the final build: small cases incl.
# a token count that does not divide over the ranks, in the cfg2 x sp2 and the sp4 layouts.
set -x
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29621"
timeout 300 $T tests/gpu_dist_check.py tiny tiny_ragged real2 > gpurun_out/dist_check_r2aa_n4.log 2>&1; grep '^{' gpurun_out/dist_check_r2aa_n4.log; tail -3 gpurun_out/dist_check_r2aa_n4.log
timeout 300 $T tests/gpu_dist_check.py cfg1 tiny tiny_ragged real2 > gpurun_out/dist_check_r2aa_n4_sp4.log 2>&1; grep '^{' gpurun_out/dist_check_r2aa_n4_sp4.log; tail -3 gpurun_out/dist_check_r2aa_n4_sp4.log
