#!/bin/bash
# Round 2, run AD (1 GPU): compute-sanitizer memcheck over the round-2 kernels at small sizes (VAE decode / encode / banded
# decode, umT5 operators + encoder, implicit-GEMM convolution, dedup / tensor-core fp32 linear, edge shapes).
set -x
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 python -m pytest tests/test_native_gpu.py -x -q -m gpu \
  -k "vae_operators or vae_decode_matches or vae_encode and tiny or row_bands or t5_operators or t5_encoder_matches and tiny or conv_as_implicit or dedup or edge_shapes" \
  > gpurun_out/sanitizer_r2ad.log 2>&1; echo rc=$?; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitizer_r2ad.log | head -20
