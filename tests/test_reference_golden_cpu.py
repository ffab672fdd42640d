"""tests/golden/reference_steps.npz was produced by running the reference's own C++ classes
(tests/golden/make_reference_fixtures.py over oracle/_ref/libaprilref.so).  On the CPU:

  * the numpy oracle reproduces every stored tensor (so the oracle is pinned by reference-generated vectors
    even where /root/reference does not exist);
  * the checker the GPU test uses (tests/golden/reference_check.py) is exercised end to end on an adapter
    over the oracle's trainer;
  * when the compiled reference is available, the fixture is regenerated in memory and must match what is
    committed (it is not stale).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from oracle import april as A  # noqa: E402
from reference_cases import CASES, build_oracle, build_reference, weight_names  # noqa: E402
from reference_check import check_trainer_against_reference, fixture, rel_err  # noqa: E402

LOSSES = {"multi_class_cross_entropy": A.MultiClassCrossEntropy, "mse": A.MSE, "cross_entropy": A.CrossEntropy}
TOL = 4e-6


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_the_reference_generated_vectors(name):
    layers, isz, osz, bunch, loss, _ = CASES[name]
    fx = fixture()
    names = weight_names(layers)
    o = build_oracle(A, layers, isz, {n: fx["%s/w/%s" % (name, n)].copy() for n in names})
    x, t = fx[name + "/x"], fx[name + "/t"]
    y = o.forward(x, True)
    L = LOSSES[loss]()
    assert rel_err(y, fx[name + "/y"]) <= TOL
    assert rel_err(L.loss_rows(y, t), fx[name + "/rows"]) <= TOL
    g = L.gradient(y, t)
    assert rel_err(g, fx[name + "/lossgrad"]) <= TOL
    assert rel_err(o.backprop(g).reshape(bunch, -1), fx[name + "/dx"]) <= TOL
    G, Cn = {}, {}
    o.compute_gradients(G, Cn)
    for n in names:
        assert rel_err(G[n], fx["%s/g/%s" % (name, n)]) <= TOL, n
        assert Cn[n] == int(fx["%s/count/%s" % (name, n)]), n


class _OracleTrainerAdapter:
    """The product trainer's method names over oracle.SupervisedTrainer."""

    def __init__(self, layers, isz, bunch, loss):
        stack = A.Stack()
        for c in build_oracle(A, layers, isz, {}).components:
            stack.push(c)
        self.tr = A.SupervisedTrainer(stack, LOSSES[loss](), bunch).build(isz)
        self.tr.set_option("learning_rate", 0.01)
        self.tr.set_option("momentum", 0.0)
        self.tr.set_option("weight_decay", 0.0)

    def set_weights(self, n, w):
        if n not in self.tr.weights or self.tr.weights[n].shape != w.shape:
            self.tr.weights[n] = w.copy()
        else:
            self.tr.weights[n][...] = w

    def calculate(self, x):
        return self.tr.calculate(x)

    def train_step(self, x, t):
        return self.tr.train_step(x, t)

    def gradients(self, n):
        return self.tr.grads[n]


@pytest.mark.parametrize("name", sorted(CASES))
def test_checker_runs_end_to_end_on_the_oracle_trainer(name):
    layers, isz, osz, bunch, loss, _ = CASES[name]
    tr = _OracleTrainerAdapter(layers, isz, bunch, loss)
    tr.calculate(fixture()[name + "/x"][:1])      # convolutions size their weights at the first forward
    check_trainer_against_reference(tr, name, weight_names(layers), bunch, TOL)


def test_committed_fixture_is_what_the_compiled_reference_produces():
    from oracle import ref as R
    if not R.available():
        pytest.skip("oracle/_ref/libaprilref.so not built (needs /root/reference)")
    fx = fixture()
    for name, (layers, isz, osz, bunch, loss, _) in CASES.items():
        net = build_reference(R, layers, isz, 0 if layers[-1][0] == "flatten" else osz)
        net.forward(np.zeros((1, isz), np.float32), False)
        net.reset(0)
        for n in weight_names(layers):
            net.set_weight(n, fx["%s/w/%s" % (name, n)])
        y = net.forward(fx[name + "/x"], True)
        assert rel_err(y, fx[name + "/y"]) <= 1e-6, name
        L = R.Loss(loss, osz)
        net.backprop(L.gradient(y, fx[name + "/t"]))
        net.compute_gradients()
        for n in weight_names(layers):
            assert rel_err(net.gradient(n), fx["%s/g/%s" % (name, n)]) <= 1e-6, (name, n)
        net.close()
