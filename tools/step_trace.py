"""Timeline of one replayed training step (CFGNAME=C1..C5, default C2) (GPU box): kernel start / duration / stream from CUPTI
through torch.profiler (the kernels are this library's; torch only hosts the profiler).

    python tools/step_trace.py [steps]
Prints the kernels of the last profiled step in start order, relative to the step's first kernel."""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import april_ann_b200 as ann  # noqa: E402
from april_ann_b200 import configs as CFG  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
torch.zeros(1, device="cuda")
ctx = ann.get_context(local_rank)
ctx.set_math_mode(ann.MATH_TF32)
name = os.environ.get("CFGNAME", "C2")
bunch = CFG.CONFIGS[name]["bunch"]
tr = CFG.build_trainer(ann, name, ctx=ctx)
tr.randomize_weights(random=ann.random(1234), inf=-1, sup=1, use_fanin=True, use_fanout=True)
if world > 1:   # replica group: python -m torch.distributed.run --nproc-per-node N tools/step_trace.py
    import torch.distributed as dist
    from april_ann_b200.parallel import init_data_parallel
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group(backend="gloo")
    init_data_parallel(tr, dist)
x, t = CFG.synthetic_bunch(name, 1)
tr.stage(x, t, bunch)
for _ in range(6):
    tr.step_staged(bunch)
ctx.sync()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        tr.step_staged(bunch)
    ctx.sync()
path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
if world > 1:
    import ctypes as C
    from april_ann_b200._lib import lib, check
    st = (C.c_longlong * 64)()
    check(lib.b200h_trainer_dp_debug(tr.h, st))
    for b in range(2):
        print("rank %d bucket %d: start %d ns after bucket 0 start; wait for peers %d, shard update %d, publish %d ns" % (
            rank, b, st[4 * b] - st[0], st[4 * b + 1] - st[4 * b], st[4 * b + 2] - st[4 * b + 1], st[4 * b + 3] - st[4 * b + 2]), flush=True)
if rank != 0:
    sys.exit(0)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
# split into steps at the first forward kernel: steps are equal-length runs
n = len(ev) // steps
last = ev[-n:]
t0 = last[0]["ts"]
print("%d kernels per step; step span %.1f us" % (n, last[-1]["ts"] + last[-1]["dur"] - t0))
for e in last:
    a = e.get("args", {})
    print("%8.1f +%6.1f  s%-3s grid %-14s %s" % (e["ts"] - t0, e["dur"], a.get("stream", "?"), str(a.get("grid", "")), e["name"][:90]))
