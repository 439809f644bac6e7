#!/bin/bash
# Round 2, run P (1 GPU): attention kernel, share of the exponentials on the FMA pipe (FX_FMHA_POLY eighths), burst and sustained.
set -x
mkdir -p gpurun_out
for P in 0 2 3 0 2; do
  FX_FMHA_POLY=$P timeout 300 python tests/gpu_microbench.py fmha 2>&1 | grep -v "^$" | sed "s/^/poly=$P /" >> gpurun_out/fmha_poly_r2p.log
done
cat gpurun_out/fmha_poly_r2p.log
