#!/bin/bash
# Round 2, run R (1 GPU): 160-wide GEMM tile + vectorised DupUp3D add: VAE parity tests, GEMM tests, encode/decode timing.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_native_gpu.py -x -q -m gpu -k "vae or gemm or conv" -s 2>&1 | tail -14 > gpurun_out/pytest_r2r.log; cat gpurun_out/pytest_r2r.log
python profiles/r2_tools/vae_once.py encode 9
python profiles/r2_tools/vae_once.py decode 3
timeout 900 python bench.py --workload vae --steps 2 --checksum-only > gpurun_out/vae_r2r.json 2> gpurun_out/vae_r2r.err; grep '^{' gpurun_out/vae_r2r.json | cut -c1-1500; tail -3 gpurun_out/vae_r2r.err
