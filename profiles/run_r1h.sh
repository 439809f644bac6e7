#!/bin/bash
# Round-1 (session 8) timing experiments on the attention softmax side (native tools only, ~1 GPU-minute):
# fmha_bexp5 = no exponentials, fmha_bexp6 = no max exchange between the two halves of a row (both give wrong results,
# only their time matters), next to the real kernel, for the shared-score-buffer pipelines.
mkdir -p gpurun_out
L=gpurun_out/fmha_softmax_exp_r1h.log
: > $L
for tool in fmha_bench fmha_bexp5 fmha_bexp6; do
  for cfg in "2 0" "2 2" "3 0" "3 2"; do
    set -- $cfg
    echo "== $tool pipe=$1 token=$2 poly=0" >> $L
    FX_FMHA_PIPE=$1 FX_FMHA_TOKEN=$2 FX_FMHA_POLY=0 timeout 120 tests/native/$tool >> $L 2>&1
  done
done
grep -v "^launch [01]" $L | cut -c1-140
