#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-2}
for be in nccl gloo; do
echo "== backend $be"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 tools/mc_probe.py $be 2>&1 | grep -v "Warning\|warn" | tail -12
done | tee gpurun_out/mc_probe.log
nvidia-smi topo -m 2>&1 | head -12
