#!/bin/bash
# Round 2, run D (8 GPUs, one box): the bench line with parity (checksum must equal the N=1 / N=2 one) and the itemised
# per-rank device-time tables, PDL off for comparison, BASELINE config 4 (50-step loop) and config 5 (long clip) in the
# cfg2 x sp4 and the sp8 (batched CFG) layouts.
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
timeout 600 $TR bench.py --gpus 8 --steps 10 --warmup 3 --itemise > gpurun_out/bench_r2d_n8.json 2> gpurun_out/bench_r2d_n8.err; tail -c 1200 gpurun_out/bench_r2d_n8.json; tail -3 gpurun_out/bench_r2d_n8.err
FX_PDL=0 timeout 400 $TR bench.py --gpus 8 --steps 10 --warmup 3 --quick > gpurun_out/bench_r2d_n8_nopdl.json 2> gpurun_out/bench_r2d_n8_nopdl.err; tail -c 400 gpurun_out/bench_r2d_n8_nopdl.json
timeout 400 $TR bench.py --gpus 8 --workload loop50 > gpurun_out/loop50_r2d_n8.json 2> gpurun_out/loop50_r2d_n8.err; cat gpurun_out/loop50_r2d_n8.json; tail -3 gpurun_out/loop50_r2d_n8.err
timeout 500 $TR bench.py --gpus 8 --workload long --steps 3 --warmup 2 --checksum-only > gpurun_out/long_r2d_n8.json 2> gpurun_out/long_r2d_n8.err; tail -c 1500 gpurun_out/long_r2d_n8.json; tail -3 gpurun_out/long_r2d_n8.err
FLEXAM_CFG_SIZE=1 timeout 500 $TR bench.py --gpus 8 --workload long --steps 3 --warmup 2 --checksum-only > gpurun_out/long_r2d_n8_sp8.json 2> gpurun_out/long_r2d_n8_sp8.err; tail -c 1500 gpurun_out/long_r2d_n8_sp8.json; tail -3 gpurun_out/long_r2d_n8_sp8.err
