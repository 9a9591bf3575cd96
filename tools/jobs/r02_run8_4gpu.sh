set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=4 --master-addr 127.0.0.1"
timeout 200 python -m pytest tests -m gpu -q -x -k "tile_modes_bit_identical or qkv_epilogue_and_joint or gemm_bias_and_gelu" 2>&1 | tail -3
TP_CASES="sp,4,272,2;sp,8,1024,1;sp,8,1024,4;sp,60,1024,1;sp,60,1024,4;p2p,60,1024,1" TP_TOL=3e-2 timeout 600 $TR --master-port 29631 tools/tp_check.py > gpurun_out/r02_sp4_cases.log 2>&1; echo rc=$?; grep -E "tp_check|Error|error" gpurun_out/r02_sp4_cases.log | tail -8
