"""BASELINE config C2 at FULL SIZE (784-2048-2048-10, bunch 1024), CUDA path against THE REFERENCE ITSELF:
the reference's own HyperplaneANNComponent / ActivationFunctionANNComponent / MultiClassCrossEntropyLossFunction
objects (oracle/_ref/libaprilref.so, compiled in place from the reference's sources; it travels to the GPU box
as a prebuilt library) run the same weights and bunch on the host cores, and the product's results are held to
them with no numpy restatement in between.

  * ReLU network (the bench's): forward log-probabilities and per-row losses.  (Gradients of a ReLU net at this
    size are compared gate-aware in test_gpu_fullsize.py: one unit within rounding of zero flips.)
  * tanh network of the same shapes: forward, losses and every smoothed weight / bias gradient.

fp32 (FFMA) mode; tolerance 1e-4 of each tensor's largest magnitude (fp32 accumulation over K = 2048 on the
device, double accumulation in the reference build's plain-loop BLAS)."""
import math
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref as R  # noqa: E402

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not R.available(), reason="oracle/_ref/libaprilref.so not present")]

BUNCH, TOL = 1024, 1e-4
NAMES = ["w1", "b1", "w2", "b2", "w3", "b3"]


@pytest.fixture(scope="module")
def ann():
    import april_ann_b200 as ann
    ann.get_context().set_math_mode(ann.MATH_FP32)
    return ann


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.abs(a - b).max() / max(1.0, np.abs(b).max()))


def reference_net(actf):
    s = R.stack()
    R.push(s, R.hyperplane(784, 2048, "w1", "b1"), R.actf(actf), R.hyperplane(2048, 2048, "w2", "b2"), R.actf(actf),
           R.hyperplane(2048, 10, "w3", "b3"), R.actf("log_softmax"))
    return R.Net(s, 784, 10)


def reference_step(net, weights, x, t):
    """forward, per-row MCCE, backward and raw gradients out of the reference's own classes"""
    for n in NAMES:
        net.set_weight(n, weights[n])
    y = net.forward(x, True)
    loss = R.Loss("multi_class_cross_entropy", 10)
    rows = loss.loss_rows(y, t)
    net.backprop(loss.gradient(y, t))
    net.compute_gradients()
    return y, rows, {n: net.gradient(n) for n in NAMES}, {n: net.shared_count(n) for n in NAMES}


def inputs():
    rs = np.random.RandomState(2026)
    x = rs.uniform(-1, 1, size=(BUNCH, 784)).astype(np.float32)
    t = np.zeros((BUNCH, 10), np.float32)
    t[np.arange(BUNCH), rs.randint(0, 10, size=BUNCH)] = 1.0
    return x, t


def product_trainer(ann, actf):
    topo = "784 inputs 2048 %s 2048 %s 10 log_softmax" % (actf, actf)
    tr = ann.trainable.supervised_trainer(ann.mlp.all_all.generate(topo), ann.loss.multi_class_cross_entropy(),
                                          BUNCH).build()
    tr.set_option("learning_rate", 0.01)
    tr.set_option("momentum", 0.0)
    tr.set_option("weight_decay", 0.0)
    tr.set_flag("keep_gradients", 1)
    tr.randomize_weights(random=ann.random(1234), inf=-1, sup=1, use_fanin=True, use_fanout=True)
    return tr


@pytest.mark.parametrize("actf", ["relu", "tanh"])
def test_c2_full_size_against_the_reference(ann, actf):
    tr = product_trainer(ann, actf)
    assert sorted(tr.weight_names()) == sorted(NAMES)
    weights = {n: tr.weights(n) for n in NAMES}
    x, t = inputs()
    net = reference_net(actf)
    y_ref, rows_ref, g_ref, counts = reference_step(net, weights, x, t)
    net.close()

    assert rel_err(tr.calculate(x), y_ref) <= TOL
    mean, rows = tr.train_step(x, t)
    assert rel_err(rows, rows_ref) <= TOL
    assert abs(mean - float(rows_ref.mean())) <= TOL * max(1.0, float(rows_ref.mean()))
    if actf == "tanh":
        for n in NAMES:
            scale = 1.0 / math.sqrt(max(counts[n], 1) * BUNCH)      # supervised.lua:797-803
            assert rel_err(tr.gradients(n), g_ref[n].astype(np.float64) * scale) <= TOL, n
