"""Runs a few replayed training steps of one BASELINE config (profiling target: ncu wraps this).
    python tools/config_step.py C4 [steps] [graph|nograph]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import april_ann_b200 as ann  # noqa: E402
from april_ann_b200 import configs as CFG  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ctx = ann.get_context(0)
ctx.set_math_mode(ann.MATH_FP32 if os.environ.get("MATH") == "fp32" else ann.MATH_TF32)
tr = CFG.build_trainer(ann, name, ctx=ctx)
if len(sys.argv) > 3 and sys.argv[3] == "nograph":
    tr.set_flag("cuda_graph", 0)
tr.randomize_weights(random=ann.random(1234), inf=-1, sup=1, use_fanin=True, use_fanout=True)
x, t = CFG.synthetic_bunch(name, 1)
bunch = CFG.CONFIGS[name]["bunch"]
tr.stage(x, t, bunch)
for _ in range(steps):
    tr.step_staged(bunch)
ctx.sync()
print("done", tr.loss_get())
