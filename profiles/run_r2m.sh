#!/bin/bash
# Round 2, run M (4 GPUs): the bench line in the cfg2 x sp2 layout on the final build (parity checksum must equal N=1/8).
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551"
timeout 600 $TR bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_r2m_n4.json 2> gpurun_out/bench_r2m_n4.err; grep '^{' gpurun_out/bench_r2m_n4.json | cut -c1-600; tail -3 gpurun_out/bench_r2m_n4.err
