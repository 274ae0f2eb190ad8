#!/bin/bash
mkdir -p gpurun_out
set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/host_profile_dist.py 256 > gpurun_out/r2c13_host_dist.txt 2>&1
head -100 gpurun_out/r2c13_host_dist.txt
