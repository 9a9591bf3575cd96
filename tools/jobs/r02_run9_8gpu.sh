set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1"
TP_CASES="sp,60,1024,1;sp,60,1024,4;p2p,60,1024,1;p2p,60,1024,4" TP_TOL=3e-2 timeout 600 $TR --master-port 29641 tools/tp_check.py > gpurun_out/r02_sp8_cases.log 2>&1; echo rc=$?; grep -E "tp_check|Error|error" gpurun_out/r02_sp8_cases.log | tail -8
