#!/bin/bash
# GPU call 18: VAE conv v2 (shared vertical taps + 256-pixel tiles), faster row kernels: parity + timing + launch list
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_vae.py -x -q 2>&1 | tail -40 > gpurun_out/r02_vae_tests_2.log
tail -12 gpurun_out/r02_vae_tests_2.log
timeout 200 python tools/vae_bench.py > gpurun_out/r02_vae_bench_2.log 2>&1
tail -4 gpurun_out/r02_vae_bench_2.log
QIMG_VAE_CONV=1 VB_SHAPES="1,128,128" timeout 100 python tools/vae_bench.py > gpurun_out/r02_vae_bench_2_v1.log 2>&1
tail -2 gpurun_out/r02_vae_bench_2_v1.log
VP_ITERS=2 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_vae_launches_2.csv python tools/vae_profile_step.py > gpurun_out/r02_vae_ncu_2.log 2>&1
wc -l gpurun_out/r02_vae_launches_2.csv
