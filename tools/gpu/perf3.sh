#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
tag=${1:-p3}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -k "tf32 or digits or train_steps or c2 or c5 or c3 or output_layer" --timeout 300 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${tag}_pytest.log
timeout 300 python tools/gemm_stamps.py 0,1,1024,2048,2048 0,1,1024,2048,2048,0,0,0,0,1 1,0,2048,784,1024 2>&1 | grep -E "^case|mma 1st|mma all|epi 1st|epi last|end  |globaltimer"
timeout 300 python tools/gemm_bench.py 0,1,1024,2048,784 0,1,1024,2048,2048 0,0,1024,2048,2048 1,0,2048,784,1024 1,0,2048,2048,1024 0,1,8192,4096,4096 2>&1 | tail -7
timeout 300 python tools/step_trace.py 4 2>&1 | tail -20
for c in C2 C3 C5; do
timeout 300 python bench.py --config $c --steps 30 --warmup 5 2>gpurun_out/${tag}_bench_$c.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('$c bench: %.0f samples/s, %.1f us/step, frac %.3f, hot %.0f, roofline %.1f frac %.3f (%.1f us) parity %s' % (d['value'], d['ms_per_step']*1e3, d['step_frac_of_tf32_peak'], d['hot_l2_value'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['launch_us'], {k:d['parity'][k] for k in ('rel_l2_w','ok')}))"
done
