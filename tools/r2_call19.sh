#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests/test_gpu_gsf.py tests/test_gpu_pipeline.py tests/test_gpu_nurbs.py tests/test_gpu_configs.py tests/test_gpu_fullsize.py -q -x > gpurun_out/r2c19_tests.log 2>&1
tail -4 gpurun_out/r2c19_tests.log
for v in "default" "TIGAR_B200_GSF_OVERLAP=0"; do
  tag=$(echo "$v" | tr ' =' '__')
  if [ "$v" = "default" ]; then v=""; fi
  env $v timeout 600 python bench.py --no-ptap --no-cpu --steps 5 --warmup 3 > gpurun_out/r2c19_bench_$tag.json 2> gpurun_out/r2c19_bench_$tag.err
  tail -2 gpurun_out/r2c19_bench_$tag.err
  python - "$tag" <<'P'
import json,sys
d=json.loads(open('gpurun_out/r2c19_bench_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], d['ms_per_step'], d['stage_ms'], 'e2e', d['e2e']['value'], d['parity']['true_relative_residual'], d['parity']['sum_U'])
for r in d['rooflines'][:4]: print("  %-50s %8.2f ms/step  hbm %.3f  fp64 %s" % (r['kernel'][:50], r['ms_per_step'], r['hbm_frac'], r['fp64_frac']))
P
done
