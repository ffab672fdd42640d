#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-2}
for args in "fp32 graph" "fp32 nograph" "tf32 graph"; do
  for fused in 1 0; do
    echo "== $args fused=$fused"
    B200_DP_FUSED=$fused timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tests/dp_debug_tool.py $args 4 2>&1 | grep -E "^step|Error|error" | head -12
  done
done > gpurun_out/dp_debug_N$N.log 2>&1
cat gpurun_out/dp_debug_N$N.log
