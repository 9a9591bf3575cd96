set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=4 --master-addr 127.0.0.1"
timeout 200 python -m pytest tests -m gpu -q -x -k "qkv_epilogue_and_joint or adversarial or large_scores or attention_backend" 2>&1 | tail -3
TP_CASES="p2p,8,1024,1;p2p,8,1024,4;p2p,60,1024,1" TP_TOL=3e-2 timeout 500 $TR --master-port 29611 tools/tp_check.py > gpurun_out/r02_tp4_cases.log 2>&1; echo rc=$?; grep tp_check gpurun_out/r02_tp4_cases.log; tail -2 gpurun_out/r02_tp4_cases.log
TP_GOLDEN=narrow_L60_H8 timeout 300 $TR --master-port 29612 tools/tp_check.py > gpurun_out/r02_tp4_golden.log 2>&1; echo rc=$?; grep tp_check gpurun_out/r02_tp4_golden.log
timeout 900 $TR --master-port 29613 bench.py --gpus 4 --steps 1 --warmup 1 > gpurun_out/r02_bench_n4_quick.json 2> gpurun_out/r02_bench_n4_quick.err; echo rc=$?; tail -c 5000 gpurun_out/r02_bench_n4_quick.json; tail -3 gpurun_out/r02_bench_n4_quick.err
