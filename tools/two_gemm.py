"""Two contractions of the C2 backward pass side by side (GPU box): the data gradient of the 2048x2048
layer on the main stream and its weight gradient on a side branch, each planned for half of the SMs,
against each of them alone.  CUPTI kernel durations through torch.profiler."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import april_ann_b200 as ann  # noqa: E402
from april_ann_b200._lib import lib, check  # noqa: E402
from april_ann_b200.ops import DeviceArray  # noqa: E402

torch.zeros(1, device="cuda")
ctx = ann.get_context(0)
ctx.set_math_mode(ann.MATH_TF32)
rng = np.random.RandomState(0)
bs, n = 1024, 2048
dY = DeviceArray.from_numpy(ctx, rng.uniform(-1, 1, (bs, n)).astype(np.float32))
W = DeviceArray.from_numpy(ctx, rng.uniform(-1, 1, (n, n)).astype(np.float32))
X = DeviceArray.from_numpy(ctx, rng.uniform(0, 1, (bs, n)).astype(np.float32))
dX = DeviceArray(ctx, (bs, n))
dW = DeviceArray(ctx, (n, n))
I = C.c_int
budget = int(sys.argv[1]) if len(sys.argv) > 1 else 74


def dgrad():
    check(lib.b200_linear_bwd_data(ctx.h, I(bs), I(n), I(n), dY.ptr, I(n), W.ptr, I(n), I(3), X.ptr, I(n), dX.ptr, I(n)))


def wgrad():
    check(lib.b200_linear_bwd_weight(ctx.h, I(bs), I(n), I(n), dY.ptr, I(n), X.ptr, I(n), C.c_float(0.03), C.c_float(0.0),
                                     dW.ptr, I(n), None))


def both(b):
    check(lib.b200_set_sm_budget(ctx.h, I(b)))
    check(lib.b200_branch_begin(ctx.h, I(2)))
    wgrad()
    check(lib.b200_branch_end(ctx.h))
    dgrad()
    check(lib.b200_set_sm_budget(ctx.h, I(0)))
    check(lib.b200_branch_join_all(ctx.h))


for _ in range(2):
    dgrad(); wgrad(); both(budget)
ctx.sync()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    dgrad(); ctx.sync()
    wgrad(); ctx.sync()
    check(lib.b200_set_sm_budget(ctx.h, I(budget))); dgrad(); ctx.sync(); wgrad(); ctx.sync(); check(lib.b200_set_sm_budget(ctx.h, I(0)))
    both(budget); ctx.sync()
ev = sorted([e for e in prof.events() if e.device_time > 0], key=lambda e: e.time_range.start)
labels = ["dgrad alone, whole device", "wgrad alone, whole device", "dgrad alone, budget %d" % budget, "wgrad alone, budget %d" % budget,
          "side by side (1)", "side by side (2)"]
t0 = ev[0].time_range.start
for e, l in zip(ev, labels):
    print("%-28s start %9.1f  dur %7.1f us  %s" % (l, e.time_range.start - t0, e.device_time, e.name[40:75]))
