#!/bin/bash
# 8 GPUs: default workload with the final code
mkdir -p gpurun_out
set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29581 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2c47_bench_256_n8.json 2> gpurun_out/r2c47_bench_256_n8.err
tail -2 gpurun_out/r2c47_bench_256_n8.err
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r2c47_bench_256_n8.json').read().strip().splitlines() if l.startswith('{')][-1])
print(d['ms_per_step'], d['stage_ms'], d['e2e']['value'], d['parity'])
P
