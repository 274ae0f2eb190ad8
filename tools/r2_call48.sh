#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_bspline.py -m gpu -q > gpurun_out/r2c48_bspline.log 2>&1
tail -3 gpurun_out/r2c48_bspline.log
