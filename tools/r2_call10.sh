#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests/test_gpu_gsf.py tests/test_gpu_pipeline.py tests/test_gpu_nurbs.py -q -x > gpurun_out/r2c10_tests.log 2>&1
tail -8 gpurun_out/r2c10_tests.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_gsf.py -q -x -k "oracle" > gpurun_out/r2c10_sanitizer.log 2>&1
tail -6 gpurun_out/r2c10_sanitizer.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_gsf.py -q -x -k "oracle" > gpurun_out/r2c10_racecheck.log 2>&1
tail -6 gpurun_out/r2c10_racecheck.log
for v in "default" "TIGAR_B200_GSF_TMA=0" "TIGAR_B200_GSF_TMA=0 TIGAR_B200_GSF_PERM=0" "TIGAR_B200_GSF_TMA=0 TIGAR_B200_GSF_PERM=0 TIGAR_B200_GSF_MINB4=1"; do
  tag=$(echo "$v" | tr ' =' '__')
  if [ "$v" = "default" ]; then v=""; fi
  env $v timeout 600 python bench.py --no-ptap --no-cpu --steps 3 --warmup 2 > gpurun_out/r2c10_bench_$tag.json 2> gpurun_out/r2c10_bench_$tag.err
  tail -2 gpurun_out/r2c10_bench_$tag.err
  python - "$tag" <<'P'
import json,sys
d=json.loads(open('gpurun_out/r2c10_bench_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], d['ms_per_step'], d['stage_ms'], d['parity']['true_relative_residual'], d['parity']['sum_U'])
for r in d['rooflines'][:5]: print("  %-50s %8.2f ms/step  hbm %.3f  fp64 %s" % (r['kernel'][:50], r['ms_per_step'], r['hbm_frac'], r['fp64_frac']))
P
done
