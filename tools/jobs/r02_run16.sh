#!/bin/bash
# GPU call 16: first run of the native VAE decode (tcgen05 TF32 implicit-GEMM convolutions): parity tests + timing
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_vae.py -x -q 2>&1 | tail -40 > gpurun_out/r02_vae_tests_1.log
cat gpurun_out/r02_vae_tests_1.log | tail -30
timeout 240 python tools/vae_bench.py > gpurun_out/r02_vae_bench_1.log 2>&1
tail -5 gpurun_out/r02_vae_bench_1.log
