#!/bin/bash
# data-parallel parity: bench.py at N ranks in three modes (fused peer-memory update in TF32 and FP32, NCCL all-reduce path)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-2}; tag=${2:-dp}; cfg=${3:-C2}
run() {  # name, env..., extra args
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 30 --warmup 5 --config $cfg $EXTRA > gpurun_out/${tag}_${name}_N$N.json 2> gpurun_out/${tag}_${name}_N$N.err
  echo "$name rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/${tag}_${name}_N$N.json") if l.startswith("{")][-1])
    print(" value %.0f ms/step %.4f e2e %.0f last_loss %.4f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["last_loss"]))
    print(" parity", {k:d["parity"][k] for k in ("rel_l2_w","worst_tensor","worst_rel_l2","loss_abs","loss_gpu","loss_oracle","replica_checksum_spread","ok")})
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/${tag}_${name}_N$N.err").read()[-2500:])
PY
}
EXTRA="" run fused_tf32 A=1
EXTRA="--math fp32" run fused_fp32 A=1
EXTRA="" run nccl_tf32 B200_DP_FUSED=0
