#!/bin/bash
mkdir -p gpurun_out
set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 8 4 2; do
timeout 600 $TR --nproc-per-node $n --master-port 2954$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2c16_bench_256_n$n.json 2> gpurun_out/r2c16_bench_256_n$n.err
tail -2 gpurun_out/r2c16_bench_256_n$n.err
done
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-ptap --no-cpu > gpurun_out/r2c16_bench_256_n1.json 2> gpurun_out/r2c16_bench_256_n1.err
python - <<'P'
import json
for n in (1,2,4,8):
    try:
        d=json.loads(open('gpurun_out/r2c16_bench_256_n%d.json'%n).read().strip().splitlines()[-1])
    except Exception as e:
        print(n, 'no json', e); continue
    print(n, d['ms_per_step'], d['stage_ms'], 'e2e', d['e2e']['value'], d['config']['cg_iterations'], d['gpu_launches'], d['parity']['sum_U'], d['parity']['true_relative_residual'])
P
