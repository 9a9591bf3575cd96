#!/bin/bash
# GPU call 22: VAE encode side (stride-2 convolutions through TMA element strides) + forward -> image + edit with pixel image
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_vae.py -q -k "stride2 or image_to_nhwc or encode or edit_request or forward_produces" 2>&1 | tail -40 > gpurun_out/r02_vae_tests_4.log
tail -30 gpurun_out/r02_vae_tests_4.log
