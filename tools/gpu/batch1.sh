#!/bin/bash
# round-2 bring-up batch 1: MMA issue rate per tile width at small grids (L2 not the bound), baseline step times of C1/C3/C4/C5
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/b1_smi.log 2>&1
timeout 300 python tools/gemm_stamps.py 0,1,2048,128,2048,128 0,1,2048,64,2048,64 0,1,2048,256,2048,256 0,1,4096,128,2048,128 0,1,8192,128,2048,128 0,1,1024,2048,2048,128 0,1,1024,2048,2048,256 0,1,1024,2048,784,128 > gpurun_out/b1_stamps.log 2>&1
timeout 300 python tools/config_bench.py C1 C3 C4 C5 > gpurun_out/b1_configs.log 2>&1
tail -5 gpurun_out/b1_configs.log
