set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02_pytest_gpu_1.log; tail -5 gpurun_out/r02_pytest_gpu_1.log
timeout 300 python tools/fmha_sweep.py > gpurun_out/r02_fmha_sweep_1.log 2>&1; cat gpurun_out/r02_fmha_sweep_1.log
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench_quick_1.json 2> gpurun_out/r02_bench_quick_1.err; tail -c 3000 gpurun_out/r02_bench_quick_1.json
