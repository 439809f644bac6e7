#!/bin/bash
# GPU call 3 (1 GPU): attention softmax scheduling experiments with the native bench (timing + accuracy, no python)
mkdir -p gpurun_out
cd tests/native
for bin in fmha_bench fmha_bench_trunc; do
 for m in sync lag; do for t in 0 1; do for p in 2 3; do
  echo "== $bin max=$m token=$t poly=$p"
  FX_FMHA_MAX=$m FX_FMHA_TOKEN=$t FX_FMHA_POLY=$p timeout 60 ./$bin | tail -2
 done; done; done
done
echo "== trace lag token=1"; FX_FMHA_MAX=lag FX_FMHA_TOKEN=1 timeout 60 ./fmha_trace | tail -18
echo "== trace sync token=1"; FX_FMHA_MAX=sync FX_FMHA_TOKEN=1 timeout 60 ./fmha_trace | tail -18
./pipe_rate_trunc | grep exp_pack
