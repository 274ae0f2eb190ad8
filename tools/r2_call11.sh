#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gsf" -c 3 -o gpurun_out/r2c11_gsf_tma \
    python bench.py --steps 1 --warmup 0 --no-ptap --no-cpu > gpurun_out/r2c11_ncu.log 2>&1
ls -la gpurun_out/r2c11_gsf_tma.ncu-rep
