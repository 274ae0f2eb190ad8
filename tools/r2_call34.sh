#!/bin/bash
# 8 GPUs: cfg4 (annulus 512^3) with the relative-error FD fit + the default workload
mkdir -p gpurun_out
set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 1200 $TR --nproc-per-node 8 --master-port 29563 bench.py --gpus 8 --workload annulus --nel 512 --steps 1 --warmup 1 > gpurun_out/r2c34_annulus_n8.json 2> gpurun_out/r2c34_annulus_n8.err
tail -3 gpurun_out/r2c34_annulus_n8.err
timeout 600 $TR --nproc-per-node 8 --master-port 29561 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2c34_bench_256_n8.json 2> gpurun_out/r2c34_bench_256_n8.err
tail -3 gpurun_out/r2c34_bench_256_n8.err
python - <<'P'
import json
for f in ['r2c34_annulus_n8','r2c34_bench_256_n8']:
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f).read().strip().splitlines() if l.startswith('{')][-1])
    except Exception as e:
        print(f, 'no json', e); continue
    print(f, d['config']['workload'], d['ms_per_step'], d['stage_ms'], d['e2e']['value'], d['config']['cg_iterations'], d['gpu_launches'])
    print(d['parity'])
P
