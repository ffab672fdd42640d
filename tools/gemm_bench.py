"""Times the tcgen05 TF32 contraction alone at the shapes of the BASELINE configs (run on the GPU
box).  CUDA events on the launching stream; `hot` = back-to-back launches (operands L2-resident
when they fit), `cold` = a 256 MiB memset between launches.

    python tools/gemm_bench.py [case ...]      case = ta,tb,M,N,K[,force_bn]
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

DEFAULT = [
    # C2 (784-2048-2048-10, bunch 1024): forward, data gradient, weight gradient
    (0, 1, 1024, 2048, 784, 0), (0, 1, 1024, 2048, 2048, 0),
    (0, 0, 1024, 2048, 2048, 0),
    (1, 0, 2048, 784, 1024, 0), (1, 0, 2048, 2048, 1024, 0),
    # C3 (4096x4096 layers, bunch 8192)
    (0, 1, 8192, 4096, 4096, 0), (0, 0, 8192, 4096, 4096, 0), (1, 0, 4096, 4096, 8192, 0),
    # C5 (512 -> 10000, bunch 4096)
    (0, 1, 4096, 10000, 512, 0), (1, 0, 10000, 512, 4096, 0),
]


def main():
    import april_ann_b200 as ann
    from april_ann_b200._lib import lib, check
    from april_ann_b200.ops import DeviceArray
    ctx = ann.get_context()
    mode = os.environ.get("GEMM_MATH", "tf32")
    ctx.set_math_mode(ann.MATH_TF32 if mode == "tf32" else ann.MATH_FP32)
    cases = DEFAULT
    if len(sys.argv) > 1:
        cases = []
        for a in sys.argv[1:]:
            v = [int(x) for x in a.split(",")]
            cases.append(tuple(v + [0] * (6 - len(v))))
    e0, e1 = C.c_void_p(), C.c_void_p()
    check(lib.b200_event_create(C.byref(e0)))
    check(lib.b200_event_create(C.byref(e1)))
    fl = DeviceArray(ctx, (64 << 20,))
    rng = np.random.RandomState(0)
    for ta, tb, M, N, K, fbn in cases:
        arr = (C.c_uint32 * 8)(*([0] * 8))
        check(lib.b200_debug_tc_override(ctx.h, C.c_int(0), arr, C.c_int(fbn)))
        A = DeviceArray.from_numpy(ctx, rng.uniform(-1, 1, (K, M) if ta else (M, K)).astype(np.float32))
        B = DeviceArray.from_numpy(ctx, rng.uniform(-1, 1, (N, K) if tb else (K, N)).astype(np.float32))
        Cm = DeviceArray(ctx, (M, N))
        lda = M if ta else K
        ldb = K if tb else N

        def launch():
            check(lib.b200_sgemm(ctx.h, C.c_int(ta), C.c_int(tb), C.c_int(M), C.c_int(N), C.c_int(K), C.c_float(1.0),
                                 A.ptr, C.c_int(lda), B.ptr, C.c_int(ldb), C.c_float(0.0), Cm.ptr, C.c_int(N)))
        for _ in range(3):
            launch()
        reps = 20
        check(lib.b200_event_record(ctx.h, e0))
        for _ in range(reps):
            launch()
        check(lib.b200_event_record(ctx.h, e1))
        ms = C.c_float()
        check(lib.b200_event_elapsed_ms(e0, e1, C.byref(ms)))
        hot = ms.value / reps * 1e3
        cold = 0.0
        for _ in range(5):
            fl.zero()
            check(lib.b200_event_record(ctx.h, e0))
            launch()
            check(lib.b200_event_record(ctx.h, e1))
            check(lib.b200_event_elapsed_ms(e0, e1, C.byref(ms)))
            cold += ms.value / 5 * 1e3
        fl_ = 2.0 * M * N * K
        print("%s ta=%d tb=%d M=%5d N=%5d K=%5d bn=%3d : hot %8.1f us %7.1f TF/s | cold %8.1f us %7.1f TF/s" % (
            mode, ta, tb, M, N, K, fbn, hot, fl_ / hot / 1e6, cold, fl_ / cold / 1e6), flush=True)
        A.free(); B.free(); Cm.free()


if __name__ == "__main__":
    main()
