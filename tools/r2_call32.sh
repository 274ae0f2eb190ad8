#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests/test_gpu_solvers.py tests/test_gpu_nurbs.py tests/test_zz_gpu_multifield.py tests/test_gpu_multi.py -q -s -m gpu > gpurun_out/r2c32_tests.log 2>&1
grep -E "passed|failed|annulus|its" gpurun_out/r2c32_tests.log | tail -12
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 2 --master-port 29551 bench.py --gpus 2 --workload annulus --nel 256 --steps 2 --warmup 1 > gpurun_out/r2c32_annulus_256_n2.json 2> gpurun_out/r2c32_annulus_256_n2.err
tail -3 gpurun_out/r2c32_annulus_256_n2.err
timeout 600 python bench.py --workload annulus --nel 128 --steps 2 --warmup 1 --no-ptap --no-cpu > gpurun_out/r2c32_annulus_128_n1.json 2> gpurun_out/r2c32_annulus_128_n1.err
tail -3 gpurun_out/r2c32_annulus_128_n1.err
timeout 600 python bench.py --steps 2 --warmup 1 --no-ptap --no-cpu > gpurun_out/r2c32_bench.json 2> gpurun_out/r2c32_bench.err
python - <<'P'
import json
for f in ['r2c32_annulus_256_n2','r2c32_annulus_128_n1','r2c32_bench']:
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f).read().strip().splitlines() if l.startswith('{')][-1])
    except Exception as e:
        print(f, 'no json', e); continue
    print(f, d['config']['workload'], d['ms_per_step'], d['stage_ms'], d['e2e']['value'], d['config']['cg_iterations'], d['gpu_launches'])
    print(d['parity'])
P
