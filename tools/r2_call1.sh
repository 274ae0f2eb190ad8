#!/bin/bash
# Round-2 GPU call 1: everything round 1 could only check on the CPU.
mkdir -p gpurun_out
set -x
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r2c1_smi.txt
TIGAR_B200_UNVERIFIED=1 timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2c1_gpu_tests.log 2>&1
timeout 300 python bench.py --mode matfree --nel 128 --steps 1 --warmup 1 --no-ptap --no-cpu \
    > gpurun_out/r2c1_matfree_128.json 2> gpurun_out/r2c1_matfree_128.err
TIGAR_B200_MF_FUSED=0 timeout 300 python bench.py --mode matfree --nel 128 --steps 1 --warmup 1 --no-ptap --no-cpu \
    > gpurun_out/r2c1_matfree2k_128.json 2> gpurun_out/r2c1_matfree2k_128.err
timeout 300 python bench.py --mode fused --nel 128 --steps 1 --warmup 1 --no-ptap --no-cpu \
    > gpurun_out/r2c1_fused_128.json 2> gpurun_out/r2c1_fused_128.err
timeout 900 python bench.py --mode matfree --nel 512 --steps 1 --warmup 0 --no-ptap --no-cpu \
    > gpurun_out/r2c1_matfree_512.json 2> gpurun_out/r2c1_matfree_512.err
tail -5 gpurun_out/r2c1_gpu_tests.log
tail -c 600 gpurun_out/r2c1_matfree_128.json gpurun_out/r2c1_matfree_512.json gpurun_out/r2c1_matfree_512.err
