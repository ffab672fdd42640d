#!/bin/bash
# bench.py on every BASELINE config at N=1 (+ the reference arm on C2); logs to gpurun_out/
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
tag=${1:-b}
for c in ${2:-C2 C1 C3 C4 C5}; do
  timeout 600 python bench.py --config $c --steps ${3:-30} --warmup 5 > gpurun_out/${tag}_bench_$c.json 2> gpurun_out/${tag}_bench_$c.err
  echo "$c rc=$?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${tag}_bench_$c.json"))
    print({k:d[k] for k in ("value","ms_per_step","step_frac_of_tf32_peak","gpu_launches")}, "e2e", d["e2e"]["value"])
    print(" roofline", {k:d["roofline"][k] for k in ("bound","achieved","peak","frac","launch_us")})
    print(" parity", d["parity"] and {k:d["parity"][k] for k in ("rel_l2_w","worst_tensor","worst_rel_l2","loss_abs","ok","oracle_seconds")})
    print(" cpu", d.get("cpu_baseline",{}).get("value"), d["clocks"])
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/${tag}_bench_$c.err").read()[-1500:])
PY
done
