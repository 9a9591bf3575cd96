set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
QIMG_FMHA_P_PARTS=4 timeout 200 python -m pytest tests -m gpu -q -x -k "qkv_epilogue_and_joint or adversarial or large_scores or headline" 2>&1 | tail -2
for parts in 2 4; do QIMG_FMHA_P_PARTS=$parts FS_MODES="6,14" FS_SHAPES="4,4224;1,4224;1,16512;4,1152" timeout 200 python tools/fmha_sweep.py 2>&1 | grep qimg | sed "s/^/parts=$parts /"; done > gpurun_out/r02_fmha_sweep_4.log; cat gpurun_out/r02_fmha_sweep_4.log
