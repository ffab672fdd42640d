#!/bin/bash
# long parity run of the replica group: k steps against the oracle at N ranks, fp32 and tf32, fused update
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-8}; K=${2:-20}
for math in fp32 tf32; do
  B200_PARITY_STEPS=$K timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
      bench.py --gpus $N --steps 5 --warmup 3 --math $math > gpurun_out/dplong_${math}_N$N.json 2> gpurun_out/dplong_${math}_N$N.err
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/dplong_${math}_N$N.json") if l.startswith("{")][-1])
print("$math N=$N k=$K parity", {k:d["parity"][k] for k in ("rel_l2_w","worst_tensor","worst_rel_l2","loss_gpu","loss_oracle","replica_checksum_spread","steps","global_bunch")})
PY
done
