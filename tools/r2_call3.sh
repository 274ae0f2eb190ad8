#!/bin/bash
# Round-2 GPU call 3: re-run fixed tests, launch list + ncu captures of the new assembly kernels.
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests/test_gpu_solvers.py tests/test_gpu_pipeline.py -q -k "config2 or sell" > gpurun_out/r2c3_tests.log 2>&1
tail -15 gpurun_out/r2c3_tests.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c3_launches_256.csv \
    python bench.py --steps 1 --warmup 1 --no-ptap --no-cpu > gpurun_out/r2c3_ncu_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/r2c3_launches_256.csv > gpurun_out/r2c3_launches_256_summary.txt
cat gpurun_out/r2c3_launches_256_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gsf -s 12 -c 6 -o gpurun_out/r2c3_gsf \
    python bench.py --steps 1 --warmup 1 --no-ptap --no-cpu --nel 128 > gpurun_out/r2c3_ncu_gsf.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tigar_qp -s 2 -c 1 -o gpurun_out/r2c3_qp \
    python bench.py --steps 1 --warmup 1 --no-ptap --no-cpu --nel 128 > gpurun_out/r2c3_ncu_qp.log 2>&1
ls -la gpurun_out/*.ncu-rep
