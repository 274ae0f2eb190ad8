#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests/test_gpu_multi.py -q -s > gpurun_out/r2c44_multi.log 2>&1
grep -E "passed|failed|annulus|MGPU" gpurun_out/r2c44_multi.log | tail -6
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 2 --master-port 29571 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2c44_bench_256_n2.json 2> gpurun_out/r2c44_bench_256_n2.err
tail -2 gpurun_out/r2c44_bench_256_n2.err
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r2c44_bench_256_n2.json').read().strip().splitlines() if l.startswith('{')][-1])
print(d['ms_per_step'], d['stage_ms'], d['e2e']['value'], d['parity'])
P
