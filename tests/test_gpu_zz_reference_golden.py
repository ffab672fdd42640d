"""The CUDA path against vectors THE REFERENCE ITSELF produced (tests/golden/reference_steps.npz, written by
tests/golden/make_reference_fixtures.py from the reference's own C++ classes compiled in place; the numpy
oracle is not involved).  For each case the product's trainer gets the fixture's weights, input and target;
its forward output, per-row losses, mean loss and smoothed weight gradients must match the reference's.

fp32 (FFMA) mode, tolerance 5e-5 of the tensor's largest magnitude (summation order only); the same cases in
TF32 mode at 5e-3 for the smooth-activation networks.  All calls go through libb200ann.so (ctypes)."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from reference_cases import CASES, build_product, weight_names  # noqa: E402
from reference_check import check_trainer_against_reference, fixture  # noqa: E402


@pytest.fixture(scope="module")
def ann():
    import april_ann_b200 as ann
    ann.get_context().set_math_mode(ann.MATH_FP32)
    return ann


def make_trainer(ann, name):
    layers, isz, osz, bunch, loss, _ = CASES[name]
    loss_obj = {"multi_class_cross_entropy": ann.loss.multi_class_cross_entropy, "mse": ann.loss.mse,
                "cross_entropy": ann.loss.cross_entropy}[loss](osz)
    tr = ann.trainable.supervised_trainer(build_product(ann, layers), loss_obj, bunch).build(isz, osz)
    tr.set_option("learning_rate", 0.01)
    tr.set_option("momentum", 0.0)
    tr.set_option("weight_decay", 0.0)
    tr.set_flag("keep_gradients", 1)
    return tr, weight_names(layers), bunch


@pytest.mark.parametrize("name", sorted(CASES))
def test_fp32_path_matches_the_reference(ann, name):
    tr, names, bunch = make_trainer(ann, name)
    assert sorted(tr.weight_names()) == sorted(names)
    for n in names:
        assert tuple(tr.weights(n).shape) == tuple(fixture()["%s/w/%s" % (name, n)].shape), n
    worst = check_trainer_against_reference(tr, name, names, bunch, 5e-5)
    print(name, "worst relative error vs the reference: %.2e" % worst)


SMOOTH = ["mlp_tanh_logistic_mcce", "mlp_softmax_mse", "mlp_log_logistic_ce", "digits_mlp_mcce"]


@pytest.mark.parametrize("name", SMOOTH)
def test_tf32_path_matches_the_reference(ann, name):
    ctx = ann.get_context()
    ctx.set_math_mode(ann.MATH_TF32)
    try:
        tr, names, bunch = make_trainer(ann, name)
        worst = check_trainer_against_reference(tr, name, names, bunch, 5e-3)
        print(name, "TF32 worst relative error vs the reference: %.2e" % worst)
    finally:
        ctx.set_math_mode(ann.MATH_FP32)
