#!/bin/bash
# GPU call 19: VAE conv with BN = 96 / 192 tiles, vectorised softmax: parity + timing + launch list + one full ncu capture
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_vae.py -x -q 2>&1 | tail -40 > gpurun_out/r02_vae_tests_3.log
tail -12 gpurun_out/r02_vae_tests_3.log
timeout 200 python tools/vae_bench.py > gpurun_out/r02_vae_bench_3.log 2>&1
tail -3 gpurun_out/r02_vae_bench_3.log
VP_ITERS=2 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_vae_launches_3.csv python tools/vae_profile_step.py > gpurun_out/r02_vae_ncu_3.log 2>&1
wc -l gpurun_out/r02_vae_launches_3.csv
# one full capture of a 1024 px 96 -> 96 convolution (the 30th conv launch of a decode is in the last up block)
VP_ITERS=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv2_tf32 -s 32 -c 1 -o gpurun_out/r02_vae_conv96 -f python tools/vae_profile_step.py > gpurun_out/r02_vae_ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep 2>/dev/null | tail -2
