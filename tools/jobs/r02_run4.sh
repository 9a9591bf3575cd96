set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02_pytest_gpu_4.log; tail -6 gpurun_out/r02_pytest_gpu_4.log
# bench, default configuration (+ eager-GPU baseline, e2e through the runner)
timeout 900 python bench.py --steps 1 --warmup 1 > gpurun_out/r02_bench_n1_a.json 2> gpurun_out/r02_bench_n1_a.err; tail -c 2500 gpurun_out/r02_bench_n1_a.json
# attention with the 25 % polynomial share, in-step
QIMG_FMHA_MODE=14 timeout 600 python bench.py --steps 1 --warmup 1 --no-e2e --no-extras --no-cpu-baseline > gpurun_out/r02_bench_n1_mode14.json 2>/dev/null; tail -c 1200 gpurun_out/r02_bench_n1_mode14.json
# BASELINE configs[4] sweep
timeout 900 python bench.py --sweep > gpurun_out/r02_sweep_n1.json 2> gpurun_out/r02_sweep_n1.err; tail -c 3000 gpurun_out/r02_sweep_n1.json
# TeaCache (reference default threshold 0.2) and CUDA-graph replay at B=1
timeout 600 python bench.py --steps 1 --warmup 1 --cache tea_cache --rel-l1-thresh 0.2 --no-e2e --no-cpu-baseline > gpurun_out/r02_bench_teacache.json 2>/dev/null; tail -c 1500 gpurun_out/r02_bench_teacache.json
timeout 600 python bench.py --steps 2 --warmup 1 --batch 1 --no-e2e --no-extras --no-cpu-baseline > gpurun_out/r02_bench_b1_eager.json 2>/dev/null; tail -c 600 gpurun_out/r02_bench_b1_eager.json
timeout 600 python bench.py --steps 2 --warmup 1 --batch 1 --graph --no-e2e --no-extras --no-cpu-baseline > gpurun_out/r02_bench_b1_graph.json 2>/dev/null; tail -c 600 gpurun_out/r02_bench_b1_graph.json
# ncu --set full: one launch of each tensor-core kernel at the bench shapes; only the raw CSV pages travel back
for c in fmha qkv outproj mlpup mlpdown; do
  pat="gemm_umma2"; [ $c = fmha ] && pat="fmha_joint"
  KB_ONLY=$c timeout 200 ncu --set full --clock-control none --import-source on -k regex:$pat -s 1 -c 1 -o /tmp/prof_r2_$c python tools/kernel_bench.py > gpurun_out/ncu_r2_$c.log 2>&1
  ncu -i /tmp/prof_r2_$c.ncu-rep --page raw --csv > gpurun_out/r02_ncu_$c.raw.csv 2>/dev/null
done
cp /tmp/prof_r2_fmha.ncu-rep gpurun_out/ 2>/dev/null
ls -la gpurun_out | tail -20
