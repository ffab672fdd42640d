#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
CFGNAME=C4 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:conv_wgrad_small|col2im|im2col|conv_fwd_kernel|maxpool_bwd" -s 20 -c 6 -o gpurun_out/prof_conv_r2 -f python tools/config_step.py C4 6 > gpurun_out/ncu_conv.log 2>&1
tail -5 gpurun_out/ncu_conv.log
ls -la gpurun_out/prof_conv_r2.ncu-rep
