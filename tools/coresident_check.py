"""Does a light HBM-bound kernel (multi-tensor SGD over 4.2 M parameters) make progress while a
persistent tcgen05 contraction occupies the SMs?  (GPU box)  Launches the contraction on the main
stream and the SGD on a side branch, and prints CUPTI start / duration of both, next to the SGD alone."""
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


class SgdTensor(C.Structure):
    _fields_ = [("w", C.c_void_p), ("g", C.c_void_p), ("u", C.c_void_p), ("n", C.c_uint64), ("rows", C.c_int32),
                ("cols", C.c_int32), ("lr", C.c_float), ("momentum", C.c_float), ("weight_decay", C.c_float),
                ("l1_norm", C.c_float), ("max_norm_penalty", C.c_float), ("pad_", C.c_int32)]


torch.zeros(1, device="cuda")
ctx = ann.get_context(0)
ctx.set_math_mode(ann.MATH_TF32)
rng = np.random.RandomState(0)
M, N, K = [int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (8192, 4096, 4096))]
A = DeviceArray.from_numpy(ctx, rng.uniform(-1, 1, (M, K)).astype(np.float32))
B = DeviceArray.from_numpy(ctx, rng.uniform(-1, 1, (N, K)).astype(np.float32))
Cm = DeviceArray(ctx, (M, N))
n = 2048 * 2048
w, g, u = (DeviceArray.from_numpy(ctx, rng.uniform(-1, 1, (n,)).astype(np.float32)) for _ in range(3))
host = SgdTensor(w.ptr.value, g.ptr.value, u.ptr.value, n, 2048, 2048, 0.01, 0.9, 1e-4, 0.0, 0.0, 0)
dev = DeviceArray(ctx, (C.sizeof(SgdTensor) // 4,))
check(lib.b200_memcpy_h2d(ctx.h, dev.ptr, C.byref(host), C.c_size_t(C.sizeof(SgdTensor))))
cnt = DeviceArray(ctx, (4,))
cnt.zero()
ctx.sync()


def gemm():
    check(lib.b200_sgemm(ctx.h, C.c_int(0), C.c_int(1), C.c_int(M), C.c_int(N), C.c_int(K), C.c_float(1.0), A.ptr, C.c_int(K),
                         B.ptr, C.c_int(K), C.c_float(0.0), Cm.ptr, C.c_int(N)))


def sgd():
    check(lib.b200_sgd_multi_tensor_ex(ctx.h, C.c_int(1), dev.ptr, C.byref(host), C.c_double(1e-5), cnt.ptr, C.c_int(0)))


for _ in range(2):
    gemm()
    sgd()
ctx.sync()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    sgd()
    ctx.sync()
    # contraction on a side branch first, then (a Python call later) the SGD on the main stream
    check(lib.b200_branch_begin(ctx.h, C.c_int(1)))
    gemm()
    check(lib.b200_branch_end(ctx.h))
    sgd()
    check(lib.b200_branch_join_all(ctx.h))
    ctx.sync()
ev = sorted([e for e in prof.events() if e.device_time > 0], key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
for e in ev:
    print("%9.1f +%8.1f us  %s" % (e.time_range.start - t0, e.device_time, e.name[:70]))
