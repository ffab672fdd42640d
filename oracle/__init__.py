"""CPU oracle for the APRIL-ANN mini-batch training hot path.

TEST INFRASTRUCTURE ONLY.  This package is a numpy restatement of the
reference's CPU (MKL/ATLAS) algorithm for the path named in BASELINE.json.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the
checker or the timed CPU baseline -- never as the thing shipped.  The product
(``april_ann_b200``) never imports this package and fails loudly when its
CUDA library is missing.

Parity status: PINNED, three ways.
(1) Against the reference itself: ``oracle/ref_build/`` compiles the reference's own C++
    components / loss functions in place into ``oracle/_ref/libaprilref.so`` (``oracle/ref.py`` is its
    ctypes face) and ``tests/test_oracle_vs_reference.py`` compares forward, backprop, gradients,
    shared counts and losses of 31 networks built from both.
(2) Against vectors the compiled reference generated (``tests/golden/reference_steps.npz``,
    ``tests/test_reference_golden_cpu.py``) -- these travel to machines without the reference.
(3) ``tests/test_oracle_golden.py`` checks this oracle against the reference's own golden vectors
    (the trainer and the optimizers are Lua in the reference and are pinned here):
  * the 10-epoch digits training curve of TEST/digitos/test.lua:15-27
    (same table in packages/ann/optimizer/test/test-digits-sgd.lua:37-49),
    tolerance 1e-3 as in the reference test;
  * the digits curves of packages/ann/optimizer/test/test-digits-{l1,rmsprop,adadelta,adagrad}.lua and
    the *ConvexTest of each optimizer;
  * the closed-form loss/gradient checks of packages/ann/loss/test/test.lua:26-116;
  * the exact integer GEMM cases of packages/basics/matrix/test/test_gemm.lua;
  * the initial validation loss / first epochs of
    packages/ann/ann/test/test-convolution-digits-output.log (conv + max-pool).

Every function cites the reference file:line it restates (paths relative to
the reference checkout root).
"""
from .mtrand import MTRand  # noqa: F401
