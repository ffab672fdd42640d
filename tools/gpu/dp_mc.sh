#!/bin/bash
# replica group over symmetric memory: parity + throughput with the NVLS kernel, the P2P kernel on the same buffers, and the IPC path
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-2}; tag=${2:-mc}
run() {
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29516 \
      bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/${tag}_${name}_N$N.json 2> gpurun_out/${tag}_${name}_N$N.err
  echo "$name rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/${tag}_${name}_N$N.json") if l.startswith("{")][-1])
    print(" value %.0f us/step %.1f e2e %.0f" % (d["value"], d["ms_per_step"]*1e3, d["e2e"]["value"]), " parity", {k:d["parity"][k] for k in ("rel_l2_w","loss_abs","replica_checksum_spread","ok")})
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/${tag}_${name}_N$N.err").read()[-3000:])
PY
}
run multicast A=1
#run symm_p2p B200_DP_MULTICAST=0
run ipc_p2p B200_DP_SYMMETRIC=0
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/step_trace.py 4 2>&1 | grep -E "bucket|step span|dp_|gemm_tc|sgd" | tail -30
