#!/bin/bash
# GPU call 21 (2 GPUs): the multi-GPU tests on the final tree + the new forward -> image test
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_vae.py -q -k "forward_produces" 2>&1 | tail -15 > gpurun_out/r02_vae_e2e_test.log
tail -5 gpurun_out/r02_vae_e2e_test.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "dp2 or tp2 or sp2 or cfg_parallel or 2gpu or two_gpu" 2>&1 | tail -15 > gpurun_out/r02_pytest_2gpu_final.log
tail -6 gpurun_out/r02_pytest_2gpu_final.log
