#!/bin/bash
# Round 2, run C (2 GPUs): parity checksums at N=2 in both layouts (cfg2 and cfg1 x sp2: fused Ulysses exchange, frame-
# sharded control fuser), itemised device time of the SP layout, the sampling loop under SP, rmsnorm/dedup tests.
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 python -m pytest tests -m gpu -q -s -k "rmsnorm or dedup or fused_exchange or tiny_tok" > gpurun_out/pytest_gpu_r2c.log 2>&1; tail -8 gpurun_out/pytest_gpu_r2c.log
timeout 300 python tests/gpu_microbench.py rows > gpurun_out/micro_rows_r2c.log 2>&1; cat gpurun_out/micro_rows_r2c.log
timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 --itemise > gpurun_out/bench_r2c_cfg2.json 2> gpurun_out/bench_r2c_cfg2.err; tail -c 1500 gpurun_out/bench_r2c_cfg2.json; tail -5 gpurun_out/bench_r2c_cfg2.err
FLEXAM_CFG_SIZE=1 timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 --itemise > gpurun_out/bench_r2c_sp2.json 2> gpurun_out/bench_r2c_sp2.err; tail -c 1500 gpurun_out/bench_r2c_sp2.json; tail -5 gpurun_out/bench_r2c_sp2.err
timeout 600 python bench.py --workload loop50 --loop-steps 6 > gpurun_out/loop6_r2c_n1.json 2> gpurun_out/loop6_r2c_n1.err; cat gpurun_out/loop6_r2c_n1.json
FLEXAM_CFG_SIZE=1 timeout 600 $TR bench.py --gpus 2 --workload loop50 --loop-steps 6 > gpurun_out/loop6_r2c_sp2.json 2> gpurun_out/loop6_r2c_sp2.err; cat gpurun_out/loop6_r2c_sp2.json; tail -3 gpurun_out/loop6_r2c_sp2.err
timeout 600 $TR bench.py --gpus 2 --workload loop50 --loop-steps 6 > gpurun_out/loop6_r2c_cfg2.json 2> gpurun_out/loop6_r2c_cfg2.err; cat gpurun_out/loop6_r2c_cfg2.json; tail -3 gpurun_out/loop6_r2c_cfg2.err
