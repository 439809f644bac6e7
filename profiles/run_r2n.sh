#!/bin/bash
# Round 2, run N (2 GPUs): VAE decode split into bands of image rows. Single-GPU threaded test with the real kernels, then
# the 2-GPU decode of the metric's clip next to the one-GPU decode (checksums must match).
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_native_gpu.py -x -q -m gpu -k "vae" -s 2>&1 | tail -15 > gpurun_out/pytest_r2n.log; cat gpurun_out/pytest_r2n.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561"
timeout 600 $TR bench.py --gpus 2 --workload vae --steps 2 > gpurun_out/vae_r2n_n2.json 2> gpurun_out/vae_r2n_n2.err; grep '^{' gpurun_out/vae_r2n_n2.json | cut -c1-900; tail -5 gpurun_out/vae_r2n_n2.err
