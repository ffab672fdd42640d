"""Bring-up harness for the tcgen05 TF32 contraction (run on the GPU box).
Integer-valued operands make every product exact in TF32, so any mismatch is a layout /
descriptor bug, not rounding.  Each case runs in its own process under a timeout."""
import ctypes as C
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [(128, 16, 32), (128, 256, 32), (128, 256, 64), (128, 112, 96), (256, 512, 128), (1024, 2048, 784),
         (1000, 300, 100), (2048, 784, 1024)]
COMBOS = [(0, 1, "NT fwd   A K-major, B K-major"), (0, 0, "NN dgrad A K-major, B MN-major"),
          (1, 0, "TN wgrad A MN-major, B MN-major"), (1, 1, "TT       A MN-major, B K-major")]


def run_case(ta, tb, M, N, K, override=None, force_bn=0):
    import numpy as np
    import april_ann_b200 as ann
    from april_ann_b200 import ops
    from april_ann_b200._lib import lib, check
    ctx = ann.get_context()
    ctx.set_math_mode(ann.MATH_TF32)
    if override is not None or force_bn:
        arr = (C.c_uint32 * 8)(*(override or [0] * 8))
        check(lib.b200_debug_tc_override(ctx.h, C.c_int(1 if override is not None else 0), arr, C.c_int(force_bn)))
    rng = np.random.RandomState(M * 7 + N * 3 + K)
    A = rng.randint(-3, 4, size=(K, M) if ta else (M, K)).astype(np.float32)
    B = rng.randint(-3, 4, size=(N, K) if tb else (K, N)).astype(np.float32)
    want = (A.T if ta else A).astype(np.float64) @ (B.T if tb else B).astype(np.float64)
    n0 = ctx.launch_count()
    got = ops.sgemm(ta, tb, 1.0, A, B)
    err = np.abs(got - want)
    bad = int((err > 1e-3).sum())
    print("M=%d N=%d K=%d ta=%d tb=%d: max_err=%.4g bad=%d/%d launches=%d" % (
        M, N, K, ta, tb, err.max(), bad, err.size, ctx.launch_count() - n0), flush=True)
    if bad:
        r, c = np.argwhere(err > 1e-3)[0]
        br, bc = np.unique(np.argwhere(err > 1e-3)[:, 0]), np.unique(np.argwhere(err > 1e-3)[:, 1])
        print("   first bad at (%d,%d): got %.3f want %.3f ; bad rows %s... bad cols %s..." % (
            r, c, got[r, c], want[r, c], br[:8], bc[:8]))
        print("   bad rows: n=%d min=%d max=%d ; bad cols: n=%d min=%d max=%d ; zero-valued among bad: %d" % (
            len(br), br.min(), br.max(), len(bc), bc.min(), bc.max(), int((got[err > 1e-3] == 0).sum())))
    return bad == 0


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "case":
        ta, tb, M, N, K = [int(v) for v in sys.argv[2:7]]
        ov = [int(v) for v in sys.argv[7:15]] if len(sys.argv) >= 15 else None
        if ov is not None and not any(ov):
            ov = None
        fbn = int(sys.argv[15]) if len(sys.argv) > 15 else 0
        ok = run_case(ta, tb, M, N, K, ov, fbn)
        sys.exit(0 if ok else 1)
    results = {}
    for ta, tb, name in COMBOS:
        print("=== " + name, flush=True)
        for (M, N, K) in CASES:
            try:
                r = subprocess.run([sys.executable, __file__, "case", str(ta), str(tb), str(M), str(N), str(K)],
                                   timeout=90, capture_output=True, text=True)
                print(r.stdout.strip() + ("" if r.returncode in (0, 1) else "\n   rc=%d %s" % (r.returncode, r.stderr.strip()[-400:])), flush=True)
                results[(ta, tb, M, N, K)] = r.returncode == 0
            except subprocess.TimeoutExpired:
                print("M=%d N=%d K=%d ta=%d tb=%d: TIMEOUT (hang)" % (M, N, K, ta, tb), flush=True)
                results[(ta, tb, M, N, K)] = False
    print("PASSED %d / %d" % (sum(results.values()), len(results)))
