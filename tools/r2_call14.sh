#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c14_gpu_tests.log 2>&1
tail -5 gpurun_out/r2c14_gpu_tests.log
TIGAR_B200_PROG_CACHE=1 timeout 900 python -m pytest tests/test_zz_gpu_multifield.py tests/test_gpu_pipeline.py -m gpu -q > gpurun_out/r2c14_progcache_tests.log 2>&1
tail -4 gpurun_out/r2c14_progcache_tests.log
timeout 900 python examples/scordelis_lo.py 256 1.0 > gpurun_out/r2c14_shell_256.log 2>&1
tail -2 gpurun_out/r2c14_shell_256.log
TIGAR_B200_PROG_CACHE=1 timeout 900 python examples/scordelis_lo.py 256 1.0 > gpurun_out/r2c14_shell_256_cache.log 2>&1
tail -2 gpurun_out/r2c14_shell_256_cache.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c14_smoke.log 2>&1
tail -5 gpurun_out/r2c14_smoke.log
