set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
# attention v10 (one query tile per CTA): correctness first (short timeout: a hang must not hold the box), then speed
timeout 300 python -m pytest tests -m gpu -q -s -x -k "qkv_epilogue_and_joint_attention or attention_backend or large_scores or adversarial" > gpurun_out/r02_pytest_fmha_v10.log 2>&1; echo rc=$?; tail -4 gpurun_out/r02_pytest_fmha_v10.log
timeout 240 python tools/fmha_sweep.py > gpurun_out/r02_fmha_sweep_2.log 2>&1; echo rc=$?; cat gpurun_out/r02_fmha_sweep_2.log
# GEMM raster band height: weight re-reads from HBM vs activation band residency
for g in 8 16 32; do QIMG_GEMM_GROUP_M=$g KB_ONLY=qkv,outproj,mlpup,mlpdown timeout 120 python tools/kernel_bench.py 2>&1 | grep -v "img only" | sed "s/^/group_m=$g /" ; done > gpurun_out/r02_gemm_group_m.log; cat gpurun_out/r02_gemm_group_m.log
