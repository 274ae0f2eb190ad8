#!/bin/bash
mkdir -p gpurun_out
set -x
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py -q -s > gpurun_out/r2c5_multi.log 2>&1
tail -25 gpurun_out/r2c5_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/r2c5_bench_256_n2.json 2> gpurun_out/r2c5_bench_256_n2.err
tail -5 gpurun_out/r2c5_bench_256_n2.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2c5_bench_256_n2.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['stage_ms'], d['e2e'], d['config']['cg_iterations'], d['gpu_launches'])
print(d['parity'])
for r in d['rooflines']: print("%-50s %7.2f ms/step  hbm %.3f  fp64 %s" % (r['kernel'][:50], r['ms_per_step'], r['hbm_frac'], r['fp64_frac']))
P
