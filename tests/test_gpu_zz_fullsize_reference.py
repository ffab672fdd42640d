"""BASELINE configs at FULL SIZE, CUDA path against THE REFERENCE ITSELF: the reference's own component and loss
objects (oracle/_ref/libaprilref.so, compiled in place from the reference's sources; it travels to the GPU box
as a prebuilt library) run the same weights and bunch on the host cores, and the product's results are held to
them with no numpy restatement in between.

  C2      784-2048-2048-10 ReLU, bunch 1024 (the bench's): forward log-probabilities and per-row losses
  C2tanh  the same shapes with tanh: forward, losses and every smoothed weight / bias gradient
  C4      the convolution / max-pooling net on 1x28x28, bunch 512: forward and losses
  C5      512 -> 10 000 log_softmax, bunch 4096: forward, losses and both gradients

(Gradients of the ReLU / max-pool networks at these sizes are compared gate-aware in test_gpu_fullsize.py: a unit
within rounding of zero flips its gate.)  fp32 (FFMA) mode; tolerance 1e-4 of each tensor's largest magnitude
(fp32 accumulation on the device, double accumulation in the reference build's plain-loop BLAS)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from oracle import ref as R  # noqa: E402
from reference_configs import NAMES, SHAPES, inputs, reference_step  # noqa: E402

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not R.available(), reason="oracle/_ref/libaprilref.so not present")]

TOL = 1e-4


@pytest.fixture(scope="module")
def ann():
    import april_ann_b200 as ann
    ann.get_context().set_math_mode(ann.MATH_FP32)
    return ann


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.abs(a - b).max() / max(1.0, np.abs(b).max()))


def product_trainer(ann, name):
    from april_ann_b200 import configs
    bunch, nin, nout = SHAPES[name]
    if name == "C2tanh":
        tr = ann.trainable.supervised_trainer(ann.mlp.all_all.generate("784 inputs 2048 tanh 2048 tanh 10 log_softmax"),
                                              ann.loss.multi_class_cross_entropy(), bunch).build()
    else:
        tr = configs.build_trainer(ann, name)
    tr.set_option("learning_rate", 0.01)
    tr.set_option("momentum", 0.0)
    tr.set_option("weight_decay", 0.0)
    tr.set_flag("keep_gradients", 1)
    tr.randomize_weights(random=ann.random(1234), inf=-1, sup=1, use_fanin=True, use_fanout=True)
    return tr


@pytest.mark.parametrize("name,with_gradients", [("C2", False), ("C2tanh", True), ("C4", False), ("C5", True)])
def test_full_size_against_the_reference(ann, name, with_gradients):
    tr = product_trainer(ann, name)
    assert sorted(tr.weight_names()) == sorted(NAMES[name])
    weights = {n: tr.weights(n) for n in NAMES[name]}
    x, t = inputs(name)
    y_ref, rows_ref, g_ref, scale = reference_step(R, name, weights, x, t)

    assert rel_err(tr.calculate(x), y_ref) <= TOL
    mean, rows = tr.train_step(x, t)
    assert rel_err(rows, rows_ref) <= TOL
    assert abs(mean - float(rows_ref.mean())) <= TOL * max(1.0, float(rows_ref.mean()))
    if with_gradients:
        for n in NAMES[name]:
            assert rel_err(tr.gradients(n), g_ref[n].astype(np.float64) * scale[n]) <= TOL, n
