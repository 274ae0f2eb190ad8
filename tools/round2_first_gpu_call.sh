#!/bin/bash
# First GPU call of the next round: everything this round could only check on the CPU, in one
# `gpurun --timeout 1500 -- bash tools/round2_first_gpu_call.sh` (logs land in gpurun_out/).
# Each step is bounded by its own timeout so a hang cannot eat the box.
mkdir -p gpurun_out
set -x
# 1. the regular GPU suite (includes the ungated multi-field tests)
timeout 600 python -m pytest tests -m gpu -x -q                        > gpurun_out/r2_gpu_tests.log 2>&1
# 2. device tests that have never run: KL shell roof, matrix-free mode, FEtoIGA
TIGAR_B200_UNVERIFIED=1 timeout 600 python -m pytest tests/test_zz_gpu_multifield.py -m gpu -q \
                                                                        > gpurun_out/r2_unverified.log 2>&1
# 3. the same multi-field tests with the per-block program cache switched on
TIGAR_B200_UNVERIFIED=1 TIGAR_B200_PROG_CACHE=1 timeout 600 python -m pytest \
    tests/test_zz_gpu_multifield.py -m gpu -q                           > gpurun_out/r2_prog_cache.log 2>&1
# 4. new examples on the device
timeout 300 python examples/elasticity.py 3 16 3                        > gpurun_out/r2_ex_elasticity.log 2>&1
timeout 300 python examples/scordelis_lo.py 16 1.0                      > gpurun_out/r2_ex_scordelis.log 2>&1
# 5. matrix-free mode: cost per CG iteration where the matrix also fits (compare with fused) ...
timeout 300 python bench.py --mode matfree --nel 128 --steps 1 --warmup 1 --no-ptap --no-cpu \
                                                                        > gpurun_out/r2_bench_matfree_128.json 2> gpurun_out/r2_bench_matfree_128.err
timeout 300 python bench.py --mode fused   --nel 128 --steps 1 --warmup 1 --no-ptap --no-cpu \
                                                                        > gpurun_out/r2_bench_fused_128.json 2> gpurun_out/r2_bench_fused_128.err
TIGAR_B200_MF_FUSED=0 timeout 300 python bench.py --mode matfree --nel 128 --steps 1 --warmup 1 --no-ptap --no-cpu \
                                                                        > gpurun_out/r2_bench_matfree_twokernel_128.json 2> gpurun_out/r2_bench_matfree_twokernel_128.err
# ... and the north_star's 512^3 patch on ONE GPU (no matrix; expect minutes)
timeout 1200 python bench.py --mode matfree --nel 512 --steps 1 --warmup 0 --no-ptap --no-cpu \
                                                                        > gpurun_out/r2_bench_matfree_512.json 2> gpurun_out/r2_bench_matfree_512.err
tail -3 gpurun_out/r2_*.log
