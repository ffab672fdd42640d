#!/bin/bash
# multi-tile-per-CTA diagnostics for the tcgen05 contraction (run on the GPU box)
for args in "0 1 4096 4096 512" "0 1 4096 4096 512 0 0 0 0 0 0 0 0 256" "0 1 4096 4096 512 0 0 0 0 0 0 0 0 240" "0 1 4096 4096 512 0 0 0 0 0 0 0 0 192" "0 1 4096 4096 512 0 0 0 0 0 0 0 0 128" "0 1 4096 4096 64 0 0 0 0 0 0 0 0 256" "0 1 2048 4096 64 0 0 0 0 0 0 0 0 128" "0 1 4096 2048 32 0 0 0 0 0 0 0 0 64" "0 1 8192 4096 4096"; do
  timeout 120 python tools/tc_debug.py case $args 2>&1 | tail -4
done
