#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
CFGNAME=C4 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:conv_wgrad_small" -s 2 -c 1 -o gpurun_out/prof_wsmall_r2 -f python tools/config_step.py C4 6 > gpurun_out/ncu_conv2.log 2>&1
tail -3 gpurun_out/ncu_conv2.log
