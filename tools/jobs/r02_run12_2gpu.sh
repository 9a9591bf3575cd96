set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29661 bench.py --gpus 2 --steps 1 --warmup 1 > gpurun_out/r02_bench_n2_c.json 2> gpurun_out/r02_bench_n2_c.err; echo rc=$?; python - <<'PY'
import json
for line in open('gpurun_out/r02_bench_n2_c.json').read().splitlines():
    if line.startswith('{'):
        d=json.loads(line); print('value',d['value'],'e2e',d['e2e']['value'],'keys',[k for k in d if k in ('sp','tp','cfg_parallel','extras_error')])
        for k in ('sp','tp'):
            for b in ('b1','b4'): print(k,b,d[k][b]['speedup_vs_n1'],d[k][b]['bit_identical_to_single_gpu'])
PY
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 1 --warmup 1 > gpurun_out/r02_bench_n1_c.json 2>/dev/null; echo rc=$?; tail -c 700 gpurun_out/r02_bench_n1_c.json
