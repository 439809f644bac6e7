#!/bin/bash
# Round-1 (session 8): timelines from builds that honour setmaxnreg (no -rdc): real kernel vs no-exponential experiment.
mkdir -p gpurun_out
L=gpurun_out/fmha_trace_r1j.log
: > $L
for tool in fmha_trace fmha_exp5; do
  for cfg in "3 0" "2 0"; do
    set -- $cfg
    echo "== $tool pipe=$1 token=$2 poly=0" >> $L
    FX_FMHA_PIPE=$1 FX_FMHA_TOKEN=$2 FX_FMHA_POLY=0 timeout 120 tests/native/$tool >> $L 2>&1
  done
done
grep -v "^launch [01]" $L | cut -c1-132
