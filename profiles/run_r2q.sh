#!/bin/bash
# Round 2, run Q (1 GPU): launch lists of one VAE encode (9 frames) and one decode (3 latent frames) at the real width.
set -x
mkdir -p gpurun_out
python profiles/r2_tools/vae_once.py encode 9
python profiles/r2_tools/vae_once.py decode 3
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/vae_enc_launches_r2q.csv python profiles/r2_tools/vae_once.py encode 9 > gpurun_out/vae_enc_r2q.log 2>&1; tail -2 gpurun_out/vae_enc_r2q.log
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/vae_dec_launches_r2q.csv python profiles/r2_tools/vae_once.py decode 3 > gpurun_out/vae_dec_r2q.log 2>&1; tail -2 gpurun_out/vae_dec_r2q.log
