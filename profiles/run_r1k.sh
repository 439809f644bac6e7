#!/bin/bash
# Round-1 (session 8): split-PV issue order (FX_FMHA_PIPE=4) against pipelines 2 and 3; timings + one timeline.
mkdir -p gpurun_out
L=gpurun_out/fmha_split_r1k.log
: > $L
for cfg in "2 0" "3 0" "4 0" "4 2" "4 1" "3 2"; do
  set -- $cfg
  for len in 11648 11500 300; do
    echo "== fmha_bench L=$len pipe=$1 token=$2 poly=0" >> $L
    FX_FMHA_PIPE=$1 FX_FMHA_TOKEN=$2 FX_FMHA_POLY=0 timeout 120 tests/native/fmha_bench $len >> $L 2>&1
  done
done
echo "== fmha_trace pipe=4 token=0 poly=0" >> $L
FX_FMHA_PIPE=4 FX_FMHA_TOKEN=0 FX_FMHA_POLY=0 timeout 120 tests/native/fmha_trace >> $L 2>&1
grep -v "^launch [01]" $L | cut -c1-132
