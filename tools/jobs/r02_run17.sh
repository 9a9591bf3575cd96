#!/bin/bash
# GPU call 17: launch list of one native VAE decode (which kernels take the 16.8 ms)
mkdir -p gpurun_out
VP_ITERS=2 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_vae_launches.csv python tools/vae_profile_step.py > gpurun_out/r02_vae_ncu.log 2>&1
tail -3 gpurun_out/r02_vae_ncu.log
wc -l gpurun_out/r02_vae_launches.csv
