"""The MT19937 streams that decide a training run -- weight initialisation (`rand(sup-inf)+inf` per element in
sorted-name order), the epoch shuffle and the dropout masks -- from three implementations that must agree draw
for draw: the reference's own MTRand (compiled, oracle/_ref), the oracle's restatement (oracle/mtrand.py) and
the product's host-side generator (april_ann_b200.random, csrc/host; no device involved, so this runs on the CPU).
Reference: packages/basics/random/c_src/MersenneTwister.{h,cc}; uniformf: basics/matrix/binding/matrix_binding.h:997-1014."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import MTRand  # noqa: E402
from oracle import ref as R  # noqa: E402

SEEDS = [1234, 5678, 0, 1, 4294967295]


@pytest.fixture(scope="module")
def ann():
    import april_ann_b200 as ann
    return ann


def have_ref():
    return R.available()


@pytest.mark.parametrize("seed", SEEDS)
def test_product_generator_matches_oracle(ann, seed):
    p, o = ann.random(seed), MTRand(seed)
    for _ in range(700):          # crosses the 624-word reload
        assert p.rand() == o.rand()
    for n in (9, 1, 255, 1000, 2 ** 31, 2 ** 32 - 1):
        for _ in range(50):
            assert p.randInt(n) == o.randInt(n), n
    assert [p.randInt(3, 12) for _ in range(40)] == [o.randInt(3, 12) for _ in range(40)]
    for size in (1, 2, 10, 800):  # 800 = the digits training set (TEST/digitos/test.lua)
        assert list(p.shuffle(size)) == list(o.shuffle(size))
    assert [p.rand(2.0) for _ in range(100)] == list(o.rand_array(100, 2.0))


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/libaprilref.so not built")
@pytest.mark.parametrize("seed", SEEDS)
def test_product_and_oracle_match_the_reference_generator(ann, seed):
    r, p, o = R.Random(seed), ann.random(seed), MTRand(seed)
    for _ in range(700):
        v = r.rand()
        assert p.rand() == v and o.rand() == v
    for n in (9, 1, 255, 1000, 2 ** 31, 2 ** 32 - 1):
        for _ in range(50):
            v = r.randint(n)
            assert p.randInt(n) == v and o.randInt(n) == v, n
    for size in (1, 2, 10, 800):
        v = r.shuffle(size)
        assert list(p.shuffle(size)) == v and list(o.shuffle(size)) == v
    # uniformf(inf, sup, random): T(random->rand(sup - inf) + inf) element by element
    lo, hi = -0.1, 0.1
    want = [np.float32(r.rand_n(hi - lo) + lo) for _ in range(300)]
    got_p = [np.float32(p.rand(hi - lo) + lo) for _ in range(300)]
    got_o = [np.float32(o.rand(hi - lo) + lo) for _ in range(300)]
    assert want == got_p == got_o
