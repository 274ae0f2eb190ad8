#!/bin/bash
# Round-2 GPU call 2: new solvers (DGEMM, band Cholesky, FD-PCG), global sum-factorised assembly.
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests/test_gpu_solvers.py tests/test_gpu_gsf.py -q -x > gpurun_out/r2c2_new_tests.log 2>&1
tail -30 gpurun_out/r2c2_new_tests.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_gsf.py -q -x -k "oracle" > gpurun_out/r2c2_sanitizer_gsf.log 2>&1
tail -15 gpurun_out/r2c2_sanitizer_gsf.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_solvers.py -q -x -k "band_cholesky or dgemm" > gpurun_out/r2c2_sanitizer_band.log 2>&1
tail -15 gpurun_out/r2c2_sanitizer_band.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2c2_gpu_tests.log 2>&1
tail -15 gpurun_out/r2c2_gpu_tests.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-ptap --no-cpu > gpurun_out/r2c2_bench_256.json 2> gpurun_out/r2c2_bench_256.err
tail -c 1500 gpurun_out/r2c2_bench_256.json; tail -5 gpurun_out/r2c2_bench_256.err
TIGAR_B200_SOLVER=jacobi timeout 600 python bench.py --steps 2 --warmup 2 --no-ptap --no-cpu > gpurun_out/r2c2_bench_256_jacobi.json 2> gpurun_out/r2c2_bench_256_jacobi.err
tail -c 600 gpurun_out/r2c2_bench_256_jacobi.json
