set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 8 --sweep --sweep-batches 1,4,16 --sweep-res 512,1024,2048 > gpurun_out/r02_sweep_n8.json 2> gpurun_out/r02_sweep_n8.err; echo rc=$?; tail -c 2500 gpurun_out/r02_sweep_n8.json
