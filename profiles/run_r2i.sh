#!/bin/bash
# Round 2, run I (1 GPU): VAE encoder (operator tests incl. stride-2 convolutions, encode vs the real module's golden /
# oracle / library execution at tiny and real width), the full clip through decode + encode.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s -k "vae or conv_as" > gpurun_out/pytest_gpu_r2i.log 2>&1; tail -30 gpurun_out/pytest_gpu_r2i.log
timeout 900 python bench.py --workload vae --vae-frames 3 --steps 1 > gpurun_out/vae3_r2i.json 2> gpurun_out/vae3_r2i.err; cat gpurun_out/vae3_r2i.json; tail -5 gpurun_out/vae3_r2i.err
timeout 900 python bench.py --workload vae --steps 1 --checksum-only > gpurun_out/vae_r2i.json 2> gpurun_out/vae_r2i.err; cat gpurun_out/vae_r2i.json; tail -5 gpurun_out/vae_r2i.err
