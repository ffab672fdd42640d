#!/bin/bash
# reduce-add split-K: correctness subset + contraction timings + step timeline / bench per backward schedule
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
tag=${1:-p2}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -k "tf32 or digits or train_steps or c2 or c5 or c3 or conv" --timeout 300 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${tag}_pytest.log
timeout 300 python tools/gemm_bench.py 0,1,1024,2048,784 0,1,1024,2048,2048 0,0,1024,2048,2048 1,0,2048,784,1024 1,0,2048,2048,1024 0,1,8192,4096,4096 2>&1 | tail -7
for mode in 1 2 0; do
  echo "== B200_CONCURRENT_BWD=$mode"
  B200_CONCURRENT_BWD=$mode timeout 300 python tools/step_trace.py 4 2>&1 | tail -22
  B200_CONCURRENT_BWD=$mode timeout 300 python bench.py --steps 30 --warmup 5 2>gpurun_out/${tag}_bench_m$mode.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('bench: %.0f samples/s, %.1f us/step, frac %.3f, hot %.0f, roofline %.1f TF/s frac %.3f (%.1f us) parity %s' % (d['value'], d['ms_per_step']*1e3, d['step_frac_of_tf32_peak'], d['hot_l2_value'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['launch_us'], {k:d['parity'][k] for k in ('rel_l2_w','ok')}))"
done 2>&1 | tee gpurun_out/${tag}_perf.log
