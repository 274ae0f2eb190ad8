#!/bin/bash
mkdir -p gpurun_out
set -x
for v in 1 0 1 0; do
TIGAR_B200_GSF_PROW=$v timeout 600 python bench.py --no-ptap --no-cpu --steps 8 > gpurun_out/r2c39_bench_$v.json 2> gpurun_out/r2c39_bench_$v.err
python - <<P
import json
d=json.loads(open('gpurun_out/r2c39_bench_$v.json').read().strip().splitlines()[-1])
print("prow=$v", d['ms_per_step'], d['stage_ms'])
for r in d['rooflines'][:4]: print("  %-50s %8.2f ms/step  hbm %.3f" % (r['kernel'][:50], r['ms_per_step'], r['hbm_frac']))
P
done
