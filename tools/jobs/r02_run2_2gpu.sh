set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
nvidia-smi topo -m | head -12
export MASTER_ADDR=127.0.0.1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1"
# 1. tensor parallel: fused push GEMM + reduce/LN/all-gather kernel vs single GPU (small first: fails fast)
timeout 300 $TR --master-port 29601 tools/tp_check.py > gpurun_out/r02_tp2_small.log 2>&1; echo rc=$?; tail -5 gpurun_out/r02_tp2_small.log
TP_CASES="p2p,8,1024,1;p2p,8,1024,4;nccl,8,1024,1;p2p,60,1024,1" TP_TOL=2e-2 timeout 600 $TR --master-port 29602 tools/tp_check.py > gpurun_out/r02_tp2_cases.log 2>&1; echo rc=$?; grep tp_check gpurun_out/r02_tp2_cases.log; tail -3 gpurun_out/r02_tp2_cases.log
# 2. the multi-GPU tests + everything touched by the step-cache / guard changes
timeout 900 python -m pytest tests -m gpu -q -s -k "tp2 or dp2 or cfg_parallel or teacache or staged or adversarial or falls_back or headline or depth or L60 or bench_batch" > gpurun_out/r02_pytest_2gpu.log 2>&1; echo rc=$?; grep -E "passed|failed|error|native vs|flagged" gpurun_out/r02_pytest_2gpu.log | tail -30
# 3. bench at N=2: DP region, runner-path e2e, TP leg, CFG-parallel leg
timeout 900 $TR --master-port 29603 bench.py --gpus 2 --steps 1 --warmup 1 > gpurun_out/r02_bench_n2_quick.json 2> gpurun_out/r02_bench_n2_quick.err; echo rc=$?; tail -c 4000 gpurun_out/r02_bench_n2_quick.json; tail -5 gpurun_out/r02_bench_n2_quick.err
