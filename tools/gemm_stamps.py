"""Per-CTA phase breakdown of one tcgen05 contraction launch (clock64 stamps; GPU box only).
    python tools/gemm_stamps.py ta,tb,M,N,K[,force_bn[,stages[,dbg_flags]]] ..."""
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
ctx.set_math_mode(ann.MATH_TF32)
rng = np.random.RandomState(0)
names = ["setup", "prod 1st issue", "mma 1st full", "mma all issued", "epi 1st tfull", "epi last done", "end"]
for a in sys.argv[1:]:
    v = [int(x) for x in a.split(",")]
    ta, tb, M, N, K = v[:5]
    fbn = v[5] if len(v) > 5 else 0
    check(lib.b200_debug_tc_stages(ctx.h, C.c_int((v[6] if len(v) > 6 else 0) | ((v[7] if len(v) > 7 else 0) << 8))))
    arr = (C.c_uint32 * 8)(*([0] * 8))
    check(lib.b200_debug_tc_override(ctx.h, C.c_int(0), arr, C.c_int(fbn)))
    A = DeviceArray.from_numpy(ctx, rng.uniform(-1, 1, (K, M) if ta else (M, K)).astype(np.float32))
    B = DeviceArray.from_numpy(ctx, rng.uniform(-1, 1, (N, K) if tb else (K, N)).astype(np.float32))
    Cm = DeviceArray(ctx, (M, N))
    st = DeviceArray(ctx, (148 * 16,), np.int64)

    def launch():
        check(lib.b200_sgemm(ctx.h, C.c_int(ta), C.c_int(tb), C.c_int(M), C.c_int(N), C.c_int(K), C.c_float(1.0),
                             A.ptr, C.c_int(M if ta else K), B.ptr, C.c_int(K if tb else N), C.c_float(0.0), Cm.ptr, C.c_int(N)))
    for _ in range(3):
        launch()
    st.zero()
    check(lib.b200_debug_tc_stamps(ctx.h, st.ptr))
    launch()
    check(lib.b200_debug_tc_stamps(ctx.h, None))
    s16 = st.numpy().reshape(148, 16)
    s16 = s16[s16[:, 0] != 0]
    s = s16[:, :8]
    rel = (s[:, 1:] - s[:, :1]).astype(np.float64)
    print("case %s: %d CTAs; cycles since CTA start (median / max over CTAs)" % (a, len(s)))
    for i, n in enumerate(names):
        print("   %-16s %9.0f %9.0f" % (n, np.median(rel[:, i]), rel[:, i].max()))
    en = ["wait staging buf", "tcgen05.ld", "split-K addend", "math + st.shared", "transposed pass", "fence + TMA issue"]
    print("   epilogue warp 2, cycles summed over its chunks (median over CTAs): " +
          ", ".join("%s %.0f" % (n, np.median(s16[:, 8 + i])) for i, n in enumerate(en)))
