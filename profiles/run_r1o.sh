#!/bin/bash
# Round-1 (session 8), 4 GPUs: multi-GPU parity (CFG2 x SP2, fused peer-write exchange through the new attention
# kernel) and the bench line at N=4.
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $T tests/gpu_dist_check.py tiny tiny_ragged real2 > gpurun_out/dist_check_n4.log 2>&1; grep '^{' gpurun_out/dist_check_n4.log; tail -3 gpurun_out/dist_check_n4.log
timeout 400 $T bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; tail -c 2500 gpurun_out/bench_n4.json; tail -3 gpurun_out/bench_n4.err
