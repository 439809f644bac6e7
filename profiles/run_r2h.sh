#!/bin/bash
# Round 2, run H (1 GPU): Wan2.2 VAE decoder — operator tests, decode vs the real module's golden / oracle / library
# execution, a short full-resolution decode (5 latent frames -> 17 frames 512x896), then the whole clip.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s -k "vae" > gpurun_out/pytest_gpu_r2h.log 2>&1; tail -30 gpurun_out/pytest_gpu_r2h.log
timeout 600 python bench.py --workload vae --vae-frames 3 --steps 1 > gpurun_out/vae3_r2h.json 2> gpurun_out/vae3_r2h.err; cat gpurun_out/vae3_r2h.json; tail -5 gpurun_out/vae3_r2h.err
timeout 900 python bench.py --workload vae --steps 1 --checksum-only > gpurun_out/vae_r2h.json 2> gpurun_out/vae_r2h.err; cat gpurun_out/vae_r2h.json; tail -5 gpurun_out/vae_r2h.err
