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
    st = DeviceArray(ctx, (148 * 32,), np.int64)

    op = v[8] if len(v) > 8 else 0
    beta = float(v[9]) if len(v) > 9 else 0.0
    Yp = DeviceArray.from_numpy(ctx, rng.uniform(-1, 1, (M, N)).astype(np.float32)) if op else None

    def launch():
        if op:   # data gradient with the relu derivative of the layer below: dX[M,N] = A[M,K] . B[K,N] (.) relu'(Yp)
            check(lib.b200_linear_bwd_data(ctx.h, C.c_int(M), C.c_int(K), C.c_int(N), A.ptr, C.c_int(K), B.ptr, C.c_int(N),
                                           C.c_int(3), Yp.ptr, C.c_int(N), Cm.ptr, C.c_int(N)))
            return
        check(lib.b200_sgemm(ctx.h, C.c_int(ta), C.c_int(tb), C.c_int(M), C.c_int(N), C.c_int(K), C.c_float(1.0),
                             A.ptr, C.c_int(M if ta else K), B.ptr, C.c_int(K if tb else N), C.c_float(beta), Cm.ptr, C.c_int(N)))
    for _ in range(3):
        launch()
    st.zero()
    check(lib.b200_debug_tc_stamps(ctx.h, st.ptr))
    if os.environ.get("PROFILE"):
        # CUPTI's view of the same launch (torch only hosts the profiler)
        import torch
        from torch.profiler import ProfilerActivity, profile
        torch.zeros(1, device="cuda")
        ctx.sync()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            launch()
            ctx.sync()
        for e in prof.events():
            if "gemm_tc" in e.name:
                print("   CUPTI kernel duration %.2f us" % e.device_time)
    else:
        launch()
    check(lib.b200_debug_tc_stamps(ctx.h, None))
    s16 = st.numpy().reshape(148, 32)
    s16 = s16[s16[:, 0] != 0]
    s = s16[:, :8]
    rel = (s[:, 1:] - s[:, :1]).astype(np.float64)
    print("case %s: %d CTAs; cycles since CTA start (median / max over CTAs)" % (a, len(s)))
    for i, n in enumerate(names):
        print("   %-16s %9.0f %9.0f" % (n, np.median(rel[:, i]), rel[:, i].max()))
    print("   globaltimer: first CTA start -> last CTA end %.2f us; CTA starts spread over %.2f us; CTA lifetime median %.2f us" % (
        (s16[:, 15].max() - s16[:, 14].min()) / 1e3, (s16[:, 14].max() - s16[:, 14].min()) / 1e3, np.median(s16[:, 15] - s16[:, 14]) / 1e3))
    if s16[:, 16].any():
        x = (s16[:, 16:19] - s16[:, :1]).astype(np.float64)
        print("   split-K exchange (warp 2): peer ready %.0f, partials sent %.0f, peer's partials here %.0f" % tuple(np.median(x, axis=0)))
    ch = (s16[:, 19:27] - s16[:, :1]).astype(np.float64)
    ch = np.where(s16[:, 19:27] != 0, ch, np.nan)
    print("   warp 2 chunk starts: " + " ".join("%.0f" % v for v in np.nanmedian(ch, axis=0) if not np.isnan(v)))
    lastc = np.nanmax(ch, axis=1, keepdims=True)
    mk = (s16[:, 27:32] - s16[:, :1]).astype(np.float64) - lastc
    print("   last chunk of warp 2, cycles since its start: after tcgen05.ld %.0f, after math+st.shared %.0f, before fence %.0f, after fence+syncwarp %.0f, after TMA issue %.0f" % tuple(np.nanmedian(mk, axis=0)))
    en = ["wait staging buf", "tcgen05.ld", "split-K addend", "math + st.shared", "transposed pass", "fence + TMA issue"]
    print("   epilogue warp 2, cycles summed over its chunks (median over CTAs): " +
          ", ".join("%s %.0f" % (n, np.median(s16[:, 8 + i])) for i, n in enumerate(en)))
