#!/bin/bash
mkdir -p gpurun_out
set -x
for v in ns4 ns6 ns8; do
if [ $v = ns4 ]; then unset TIGAR_B200_LIB; else export TIGAR_B200_LIB=$PWD/gpurun_variants/libtigar_$v.so; fi
timeout 600 python bench.py --no-ptap --no-cpu > gpurun_out/r2c30_bench_$v.json 2> gpurun_out/r2c30_bench_$v.err
tail -2 gpurun_out/r2c30_bench_$v.err
python - <<P
import json
d=json.loads(open('gpurun_out/r2c30_bench_$v.json').read().strip().splitlines()[-1])
print("$v", d['ms_per_step'], d['stage_ms'], d['e2e']['value'], d['parity']['sum_U'], d['parity']['true_relative_residual'])
for r in d['rooflines'][:4]: print("  %-50s %8.2f ms/step  hbm %.3f  fp64 %s" % (r['kernel'][:50], r['ms_per_step'], r['hbm_frac'], r['fp64_frac']))
P
done
