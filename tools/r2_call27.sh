#!/bin/bash
# 8 GPUs of one node: strong scaling of the default workload and BASELINE configs[3] (annulus)
mkdir -p gpurun_out
set -x
nvidia-smi -L | wc -l
grep -E "MemTotal|MemAvailable" /proc/meminfo
AVAIL=$(awk '/MemAvailable/ {print int($2/1048576)}' /proc/meminfo)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2c27_bench_256_n8.json 2> gpurun_out/r2c27_bench_256_n8.err
tail -3 gpurun_out/r2c27_bench_256_n8.err
timeout 600 $TR --nproc-per-node 4 --master-port 29542 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2c27_bench_256_n4.json 2> gpurun_out/r2c27_bench_256_n4.err
tail -3 gpurun_out/r2c27_bench_256_n4.err
NEL=512; if [ "$AVAIL" -lt 300 ]; then NEL=384; fi
timeout 1200 $TR --nproc-per-node 8 --master-port 29543 bench.py --gpus 8 --workload annulus --nel $NEL --steps 1 --warmup 1 > gpurun_out/r2c27_annulus_n8.json 2> gpurun_out/r2c27_annulus_n8.err
tail -5 gpurun_out/r2c27_annulus_n8.err
python - <<'P'
import json
for f in ['r2c27_bench_256_n8','r2c27_bench_256_n4','r2c27_annulus_n8']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'no json', e); continue
    print(f, d['config']['workload'], d['ms_per_step'], d['stage_ms'], d['e2e']['value'], d['config']['cg_iterations'], d['gpu_launches'])
    print(d['parity'])
    for r in d['rooflines'][:6]: print("  %-50s %8.2f ms/step  hbm %.3f  fp64 %s" % (r['kernel'][:50], r['ms_per_step'], r['hbm_frac'], r['fp64_frac']))
P
