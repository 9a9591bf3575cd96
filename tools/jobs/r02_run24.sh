#!/bin/bash
# GPU call 24: final state of the round on one B200: VAE decode + encode timing, the whole GPU suite, the bench line
mkdir -p gpurun_out
timeout 240 python tools/vae_bench.py > gpurun_out/r02_vae_bench_5.log 2>&1
tail -4 gpurun_out/r02_vae_bench_5.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r02_gpu_tests_final.log
tail -6 gpurun_out/r02_gpu_tests_final.log
timeout 600 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_final.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['whole_step']['frac'], d['clocks'], d.get('vae_decode',{}).get('native_ms'))
PY
