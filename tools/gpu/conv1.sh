#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
tag=${1:-c1}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_features.py -q -k "conv or c4 or scratch or tf32" --timeout 300 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/${tag}_pytest.log
CFGNAME=C4 timeout 300 python tools/step_trace.py 4 2>&1 | tail -60
timeout 300 python bench.py --config C4 --steps 30 --warmup 5 2>gpurun_out/${tag}_bench_C4.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('C4 bench: %.0f samples/s, %.1f us/step, hot %.0f, roofline %.1f frac %.3f (%.1f us) parity %s' % (d['value'], d['ms_per_step']*1e3, d['hot_l2_value'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['launch_us'], {k:d['parity'][k] for k in ('rel_l2_w','ok')}))"
tail -3 gpurun_out/${tag}_bench_C4.err
