#!/bin/bash
# GPU call 20: final validation of the round: VAE (deeper weight ring at BN = 96) + smoke + the whole GPU suite + bench line
mkdir -p gpurun_out
timeout 200 python tools/vae_bench.py > gpurun_out/r02_vae_bench_4.log 2>&1
tail -3 gpurun_out/r02_vae_bench_4.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_final.log 2>&1
tail -3 gpurun_out/r02_smoke_final.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r02_gpu_tests_final.log
tail -8 gpurun_out/r02_gpu_tests_final.log
timeout 600 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
tail -c 3000 gpurun_out/r02_bench_final.json
