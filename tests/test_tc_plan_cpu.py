"""Host-side launch plan of the tcgen05 contraction (csrc/gemm_tc.cu: pick_tile + finish_plan), swept over
shapes on the CPU through b200_debug_tc_plan -- the same two functions gemm_tc() calls before it encodes
tensor maps and launches.  A wrong plan is a wrong result or a hang on the GPU (an empty k slice adds an
uninitialised accumulator; a ring that does not fit shared memory fails the launch), so the invariants the
kernel relies on are checked here for every BASELINE shape and a few thousand random ones."""
import ctypes as C
import os
import random

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BM, BK = 128, 32
SMEM_MAX = 227 * 1024
A_BYTES = BM * BK * 4


@pytest.fixture(scope="module")
def plan():
    lib = C.CDLL(os.path.join(ROOT, "april_ann_b200", "libb200ann.so"))
    fn = lib.b200_debug_tc_plan
    fn.argtypes = [C.c_int] * 6 + [C.POINTER(C.c_int)] * 5

    def call(M, N, K, b_kmajor=1, sms=148, reduce_add=0):
        out = [C.c_int() for _ in range(5)]
        assert fn(M, N, K, b_kmajor, sms, reduce_add, *[C.byref(o) for o in out]) == 0
        return dict(zip(("bn", "splitk", "stages", "stage_bytes", "staging_bytes"), [o.value for o in out]))
    return call


def check(p, M, N, K, b_kmajor, sms, reduce_add):
    nk = -(-K // BK)
    # tile width: a multiple of 32 up to 256, or 16 for very narrow outputs
    assert p["bn"] == 16 if N <= 16 else (p["bn"] % 32 == 0 and 32 <= p["bn"] <= 256), p
    # k slices: a power of two; exchange mode pairs only; every slice owns at least one k-block
    s = p["splitk"]
    assert s >= 1 and (s & (s - 1)) == 0 and s <= (64 if reduce_add else 2), p
    per = -(-nk // s)
    assert (s - 1) * per < nk, ("empty k slice", p, nk)
    # ring: at least two stages, at most eight, and the whole thing fits the SM's shared memory
    b_bytes = p["bn"] * BK * 4 if b_kmajor else -(-p["bn"] // 32) * 4096
    assert p["stage_bytes"] == A_BYTES + b_bytes and p["stage_bytes"] % 1024 == 0, p
    assert 2 <= p["stages"] <= 8, p
    # ring + staging tiles + barriers (256 B) + bias row (1 KiB) + alignment slack (1 KiB): gemm_tc_kernel.cuh SMEM_EXTRA
    assert p["stages"] * p["stage_bytes"] + p["staging_bytes"] + 256 + 1024 + 1024 <= SMEM_MAX, p
    assert p["staging_bytes"] in (4 * 4096, 8 * 4096), p
    # exchange-mode pairs must be co-resident: two CTAs per tile in one wave
    tiles = -(-M // BM) * -(-N // p["bn"])
    if s == 2 and not reduce_add:
        assert 2 * tiles <= sms, p


BASELINE_SHAPES = [
    # (M, N, K, b_kmajor, reduce_add)       C2: fwd1, fwd2, dX2, dW2, dW1
    (1024, 2048, 784, 1, 0), (1024, 2048, 2048, 1, 0), (1024, 2048, 2048, 0, 0), (2048, 2048, 1024, 0, 1),
    (2048, 784, 1024, 0, 1),
    (8192, 4096, 4096, 1, 0), (8192, 4096, 4096, 0, 0), (4096, 4096, 8192, 0, 1),          # C3
    (4096, 10000, 512, 1, 0), (10000, 512, 4096, 0, 1),                                      # C5
    (32768, 32, 400, 1, 0), (32768, 400, 32, 0, 0), (32, 400, 32768, 0, 1),                  # C4 conv2 as im2col
    (32, 256, 256, 1, 0), (512, 256, 512, 1, 0),                                             # C1, C4 dense
]


@pytest.mark.parametrize("sms", [148, 74, 64])
def test_baseline_shapes(plan, sms):
    for M, N, K, bk, ra in BASELINE_SHAPES:
        check(plan(M, N, K, bk, sms, ra), M, N, K, bk, sms, ra)


def test_known_decisions(plan):
    # one 128x256 tile per CTA pair with split-K for the C2 forward (64 tiles on 148 SMs)
    p = plan(1024, 2048, 2048, 1, 148, 0)
    assert (p["bn"], p["splitk"]) == (256, 2)
    # C3: many more tiles than SMs -> full-width tiles, no split, persistent CTAs with 8 staging tiles
    p = plan(8192, 4096, 4096, 1, 148, 0)
    assert (p["bn"], p["splitk"], p["staging_bytes"]) == (256, 1, 8 * 4096)
    # the convolution weight gradient: two output tiles, 1024 k-blocks -> cut deep to fill the device
    p = plan(32, 400, 32768, 0, 148, 1)
    assert p["splitk"] >= 32
    # a contraction too short to split
    assert plan(1024, 2048, 64, 1, 148, 0)["splitk"] == 1


def test_random_shapes(plan):
    rnd = random.Random(20261018)
    for _ in range(4000):
        M = rnd.choice([1, 7, 32, 100, 128, 129, 512, 1024, 4096, 32768]) + rnd.randrange(0, 3)
        N = rnd.choice([1, 8, 10, 16, 17, 32, 33, 100, 256, 257, 400, 2048, 10000]) + rnd.randrange(0, 3)
        K = rnd.choice([8, 25, 31, 32, 33, 64, 127, 128, 400, 784, 2048, 4096, 32768]) + rnd.randrange(0, 3)
        bk, ra = rnd.randrange(2), rnd.randrange(2)
        sms = rnd.choice([148, 132, 74, 64, 16, 1])
        check(plan(M, N, K, bk, sms, ra), M, N, K, bk, sms, ra)
