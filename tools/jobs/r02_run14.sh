set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for g in 1 2 4 6 16; do QIMG_GEMM_GROUP_M=$g KB_ONLY=qkv,outproj,mlpup,mlpdown timeout 120 python tools/kernel_bench.py 2>&1 | grep -v "img only" | grep gemm | sed "s/^/group_m=$g /" ; done > gpurun_out/r02_gemm_group_m2.log; cat gpurun_out/r02_gemm_group_m2.log
for g in 2 6; do for c in qkv mlpup mlpdown; do
  QIMG_GEMM_GROUP_M=$g KB_ONLY=$c timeout 120 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:gemm_umma2 -s 1 -c 1 --csv python tools/kernel_bench.py 2>/dev/null | grep -E "dram__bytes|hit_rate|time_duration" | sed "s/^/g=$g $c /"
done; done > gpurun_out/r02_gemm_group_m2_dram.log; cut -c1-200 gpurun_out/r02_gemm_group_m2_dram.log
