set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r02_pytest_gpu_final.log 2>&1; echo rc=$?; grep -E "passed|failed|rror" gpurun_out/r02_pytest_gpu_final.log | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; echo rc=$?; tail -c 1500 gpurun_out/r02_bench_n1_final.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 | tail -c 600
