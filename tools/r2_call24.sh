#!/bin/bash
# 2 GPUs: slab upload of a host control net (NURBS annulus), multi-GPU parity worker, scaling leg
mkdir -p gpurun_out
set -x
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_bspline.py -q -m gpu > gpurun_out/r2c24_bspline.log 2>&1
tail -3 gpurun_out/r2c24_bspline.log
timeout 900 python -m pytest tests/test_gpu_multi.py -q -s > gpurun_out/r2c24_multi.log 2>&1
tail -12 gpurun_out/r2c24_multi.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 2 --master-port 29531 bench.py --gpus 2 --workload annulus --nel 256 --steps 2 --warmup 1 > gpurun_out/r2c24_annulus_256_n2.json 2> gpurun_out/r2c24_annulus_256_n2.err
tail -5 gpurun_out/r2c24_annulus_256_n2.err
timeout 900 $TR --nproc-per-node 2 --master-port 29532 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2c24_bench_256_n2.json 2> gpurun_out/r2c24_bench_256_n2.err
tail -5 gpurun_out/r2c24_bench_256_n2.err
python - <<'P'
import json
for f in ['r2c24_annulus_256_n2','r2c24_bench_256_n2']:
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f).read().strip().splitlines() if l.startswith('{')][-1])
    except Exception as e:
        print(f, 'no json', e); continue
    print(f, d['config']['workload'], d['ms_per_step'], d['stage_ms'], d['e2e'], d['config']['cg_iterations'], d['gpu_launches'])
    print(d['parity'])
    for r in d['rooflines'][:6]: print("  %-50s %8.2f ms/step  hbm %.3f  fp64 %s" % (r['kernel'][:50], r['ms_per_step'], r['hbm_frac'], r['fp64_frac']))
P
