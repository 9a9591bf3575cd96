#!/bin/bash
# ncu --set full capture of one launch of each hot kernel at the bench shapes (B=4, 1024px).  Run under gpurun (1 GPU).
mkdir -p gpurun_out
for c in qkv outproj mlpup mlpdown; do
  KB_ONLY=$c timeout 200 ncu --set full --clock-control none --import-source on -k regex:"gemm_umma2" -s 1 -c 1 \
     -o gpurun_out/prof_r1_gemm2_$c python tools/kernel_bench.py > gpurun_out/ncu_$c.log 2>&1
done
KB_ONLY=fmha timeout 200 ncu --set full --clock-control none --import-source on -k regex:"fmha_joint" -s 1 -c 1 \
   -o gpurun_out/prof_r1_fmha_final python tools/kernel_bench.py > gpurun_out/ncu_fmha.log 2>&1
KB_ONLY=ew timeout 200 ncu --set full --clock-control none -k regex:"ln_modulate|linear_small_m|cfg_euler" -c 6 \
   -o gpurun_out/prof_r1_ew python tools/kernel_bench.py > gpurun_out/ncu_ew.log 2>&1
# launch list of one quick bench step (shares, not absolutes)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 5600 -c 1150 --csv --log-file gpurun_out/launches_r1.csv \
   python bench.py --num-inference-steps 2 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
ls -la gpurun_out | tail -12
