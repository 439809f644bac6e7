#!/bin/bash
# Round 2, run AC (2 GPUs): torch.library operators on the GPU; the NCCL (non-fused) form of the Ulysses exchange and the fused
# one on small cases at 2 GPUs (sequence-parallel layout), final build.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_native_gpu.py -x -q -m gpu -k "torch_library or missing_extension" 2>&1 | tail -6 > gpurun_out/pytest_r2ac.log; cat gpurun_out/pytest_r2ac.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641"
FLEXAM_SP_EXCHANGE=nccl timeout 300 $T tests/gpu_dist_check.py cfg1 tiny tiny_ragged real2 > gpurun_out/dist_check_r2ac_nccl.log 2>&1; grep '^{' gpurun_out/dist_check_r2ac_nccl.log; tail -2 gpurun_out/dist_check_r2ac_nccl.log
timeout 300 $T tests/gpu_dist_check.py cfg1 tiny tiny_ragged real2 > gpurun_out/dist_check_r2ac_fused.log 2>&1; grep '^{' gpurun_out/dist_check_r2ac_fused.log; tail -2 gpurun_out/dist_check_r2ac_fused.log
