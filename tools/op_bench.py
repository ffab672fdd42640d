"""Times the HBM-bound kernels of the C2 step alone (GPU box): skinny output layer, bias
gradient, fused log_softmax+MCCE, multi-tensor SGD.  CUDA events on the launching stream; `cold`
= 256 MiB memset between launches.  Prints achieved GB/s against the algorithmic bytes."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import april_ann_b200 as ann  # noqa: E402
from april_ann_b200._lib import lib, check  # noqa: E402
from april_ann_b200.ops import DeviceArray  # noqa: E402

ctx = ann.get_context()
e0, e1 = C.c_void_p(), C.c_void_p()
check(lib.b200_event_create(C.byref(e0)))
check(lib.b200_event_create(C.byref(e1)))
fl = DeviceArray(ctx, (64 << 20,))
rng = np.random.RandomState(0)


def dev(*shape, lo=-1.0, hi=1.0):
    return DeviceArray.from_numpy(ctx, rng.uniform(lo, hi, shape).astype(np.float32))


def timeit(name, fn, bytes_):
    for _ in range(3):
        fn()
    ms = C.c_float()
    check(lib.b200_event_record(ctx.h, e0))
    for _ in range(20):
        fn()
    check(lib.b200_event_record(ctx.h, e1))
    check(lib.b200_event_elapsed_ms(e0, e1, C.byref(ms)))
    hot = ms.value / 20 * 1e3
    cold = 0.0
    for _ in range(5):
        fl.zero()
        check(lib.b200_event_record(ctx.h, e0))
        fn()
        check(lib.b200_event_record(ctx.h, e1))
        check(lib.b200_event_elapsed_ms(e0, e1, C.byref(ms)))
        cold += ms.value / 5 * 1e3
    print("%-44s hot %7.1f us %7.0f GB/s | cold %7.1f us %7.0f GB/s" % (name, hot, bytes_ / hot / 1e3, cold, bytes_ / cold / 1e3), flush=True)


I = C.c_int
M, K, N = 1024, 2048, 10
if len(sys.argv) > 3:
    M, K, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
X, W, b = dev(M, K), dev(N, K, lo=-.1, hi=.1), dev(N)
Y, dY, dX, Yp = DeviceArray(ctx, (M, N)), dev(M, N), DeviceArray(ctx, (M, K)), dev(M, K, lo=0, hi=1)
dW, db = DeviceArray(ctx, (N, K)), DeviceArray(ctx, (N,))
timeit("skinny fwd  %dx%dx%d" % (M, N, K), lambda: check(lib.b200_linear_fwd(ctx.h, I(M), I(N), I(K), X.ptr, I(K), W.ptr, I(K), b.ptr, I(0), Y.ptr, I(N))), 4.0 * (M * K + N * K + M * N))
timeit("skinny dgrad (+relu') %dx%dx%d" % (M, N, K), lambda: check(lib.b200_linear_bwd_data(ctx.h, I(M), I(N), I(K), dY.ptr, I(N), W.ptr, I(K), I(3), Yp.ptr, I(K), dX.ptr, I(K))), 4.0 * (2 * M * K + N * K + M * N))
timeit("skinny wgrad (+db) %dx%dx%d" % (M, N, K), lambda: check(lib.b200_linear_bwd_weight(ctx.h, I(M), I(N), I(K), dY.ptr, I(N), X.ptr, I(K), C.c_float(0.03), C.c_float(0.0), dW.ptr, I(K), db.ptr)), 4.0 * (M * K + N * K + M * N))
NB = 2048
dYb, dbb = dev(M, NB), DeviceArray(ctx, (NB,))
timeit("bias grad %dx%d" % (M, NB), lambda: check(lib.b200_bias_grad(ctx.h, I(M), I(NB), dYb.ptr, I(NB), C.c_float(0.03), C.c_float(0.0), dbb.ptr)), 4.0 * (M * NB + NB))
for Cc in (10, 10000):
    Mc = 1024 if Cc == 10 else 4096
    z, t = dev(Mc, Cc, lo=-4, hi=4), DeviceArray(ctx, (Mc, Cc))
    t.zero()
    lp, rows, g = DeviceArray(ctx, (Mc, Cc)), DeviceArray(ctx, (Mc,)), DeviceArray(ctx, (Mc, Cc))
    timeit("log_softmax+MCCE+grad %dx%d" % (Mc, Cc), lambda: check(lib.b200_log_softmax_mcce_fused(ctx.h, I(Mc), I(Cc), z.ptr, t.ptr, lp.ptr, rows.ptr, g.ptr)), 16.0 * Mc * Cc)
