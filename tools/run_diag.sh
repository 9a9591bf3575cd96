#!/bin/bash
# one python process per check, each under its own timeout so a hung kernel cannot eat the GPU lease
mkdir -p gpurun_out
for c in probe elementwise gemm qkv_fmha model; do
  echo "##### $c" >> gpurun_out/diag.log
  timeout 300 python tools/gpu_diag.py $c >> gpurun_out/diag.log 2>&1
  echo "exit=$?" >> gpurun_out/diag.log
done
tail -c 6000 gpurun_out/diag.log
