#!/bin/bash
# Round-1 (session 8) follow-up, one short gpurun call, native tools only: timelines of BOTH query tiles for the
# shared-score-buffer attention pipeline under the three turn-taking modes (FX_FMHA_TOKEN=0 none, 1 hand-over after
# the exponential pass, 2 hand-over after its first half), plus timings of the combinations.
mkdir -p gpurun_out
L=gpurun_out/fmha_token_r1g.log
: > $L
for cfg in "2 1 0" "2 2 0" "2 0 0" "2 2 2" "3 2 0" "3 1 0" "1 1 2"; do
  set -- $cfg
  echo "== fmha_bench pipe=$1 token=$2 poly=$3" >> $L
  FX_FMHA_PIPE=$1 FX_FMHA_TOKEN=$2 FX_FMHA_POLY=$3 timeout 120 tests/native/fmha_bench >> $L 2>&1
done
for cfg in "2 1 0" "2 2 0" "2 0 0" "3 2 0"; do
  set -- $cfg
  echo "== fmha_trace pipe=$1 token=$2 poly=$3" >> $L
  FX_FMHA_PIPE=$1 FX_FMHA_TOKEN=$2 FX_FMHA_POLY=$3 timeout 120 tests/native/fmha_trace >> $L 2>&1
done
grep -v "^launch [01]" $L | cut -c1-140
