#!/bin/bash
# Round-1 (session 8): MMA side + synchronisation alone (fmha_bexp7) and with the TMEM ld/st traffic (fmha_bexp8).
mkdir -p gpurun_out
L=gpurun_out/fmha_softmax_exp_r1i.log
: > $L
for tool in fmha_bexp7 fmha_bexp8; do
  for cfg in "2 0" "3 0"; do
    set -- $cfg
    for wa in 0 1; do
      echo "== $tool pipe=$1 token=$2 warp_arrive=$wa poly=0" >> $L
      FX_FMHA_PIPE=$1 FX_FMHA_TOKEN=$2 FX_FMHA_WARP_ARRIVE=$wa FX_FMHA_POLY=0 timeout 120 tests/native/$tool >> $L 2>&1
    done
  done
done
grep -v "^launch [01]" $L | cut -c1-140
