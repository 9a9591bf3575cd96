set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1"
TP_CASES="sp,2,512,1;sp,3,272,2" timeout 300 $TR --master-port 29621 tools/tp_check.py > gpurun_out/r02_sp2_small.log 2>&1; echo rc=$?; grep -E "tp_check|Error|error" gpurun_out/r02_sp2_small.log | tail -8
TP_CASES="sp,8,1024,1;sp,8,1024,4;sp,60,1024,1;p2p,8,1024,1" TP_TOL=2e-2 timeout 600 $TR --master-port 29622 tools/tp_check.py > gpurun_out/r02_sp2_cases.log 2>&1; echo rc=$?; grep -E "tp_check|Error|error" gpurun_out/r02_sp2_cases.log | tail -8
timeout 600 python -m pytest tests -m gpu -q -s -k "sp2 or tp2 or dp2 or cfg_parallel or cross_request or edit_plus" > gpurun_out/r02_pytest_2gpu_b.log 2>&1; echo rc=$?; grep -E "passed|failed|rror" gpurun_out/r02_pytest_2gpu_b.log | tail -8
timeout 900 $TR --master-port 29623 bench.py --gpus 2 --steps 1 --warmup 1 > gpurun_out/r02_bench_n2_b.json 2> gpurun_out/r02_bench_n2_b.err; echo rc=$?; tail -c 6000 gpurun_out/r02_bench_n2_b.json | cut -c1-6000; tail -3 gpurun_out/r02_bench_n2_b.err
