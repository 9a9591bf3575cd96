#!/bin/bash
# GPU call 23: AdaLayerNorm per-token modulation index (kernel + gather + plug-in) against the reference layer's golden
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "adalayernorm or ln_modulate" 2>&1 | tail -15 > gpurun_out/r02_adaln_index_tests.log
tail -8 gpurun_out/r02_adaln_index_tests.log
