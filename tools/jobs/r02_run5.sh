set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
FS_MODES="6,7,15" FS_SHAPES="4,4224;1,4224;1,16512;2,300" timeout 240 python tools/fmha_sweep.py > gpurun_out/r02_fmha_sweep_3.log 2>&1; echo rc=$?; cat gpurun_out/r02_fmha_sweep_3.log
# MLP-down GEMM with the per-problem band height (8 tiles at K = 12288): DRAM traffic
KB_ONLY=mlpdown timeout 200 ncu --set full --clock-control none -k regex:gemm_umma2 -s 1 -c 1 -o /tmp/prof_r2_mlpdown_b python tools/kernel_bench.py > gpurun_out/ncu_r2_mlpdown_b.log 2>&1
ncu -i /tmp/prof_r2_mlpdown_b.ncu-rep --page raw --csv > gpurun_out/r02_ncu_mlpdown_b.raw.csv 2>/dev/null
grep -E "best" gpurun_out/ncu_r2_mlpdown_b.log | head -3
