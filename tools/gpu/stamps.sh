#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/gemm_stamps.py "$@" > gpurun_out/stamps.log 2>&1
grep -E "^case|mma 1st|mma all|epi 1st|epi last|end  |globaltimer|split-K" gpurun_out/stamps.log
