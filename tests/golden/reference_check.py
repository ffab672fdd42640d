"""The comparison every consumer of tests/golden/reference_steps.npz runs: a trainer-like object (the product's
supervised_trainer on the GPU, or an adapter over the numpy oracle on the CPU) against what the reference's
own classes produced for the same weights, input and target."""
import math
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_FIX = None


def fixture():
    global _FIX
    if _FIX is None:
        _FIX = np.load(os.path.join(HERE, "reference_steps.npz"))
    return _FIX


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.abs(a - b).max() / max(1.0, np.abs(b).max())) if b.size else 0.0


def check_trainer_against_reference(tr, name, names, bunch, tol):
    """tr: .set_weights(n, w) .calculate(x) .train_step(x, t) -> (mean, rows) .gradients(n), built for case `name`
    with smoothing on, weight decay 0 and gradients kept.  Returns the worst relative error seen."""
    fx = fixture()
    for n in names:
        tr.set_weights(n, fx["%s/w/%s" % (name, n)])
    x, t = fx[name + "/x"], fx[name + "/t"]
    worst = {}
    worst["forward"] = rel_err(np.asarray(tr.calculate(x)).reshape(fx[name + "/y"].shape), fx[name + "/y"])
    mean, rows = tr.train_step(x, t)
    worst["loss_rows"] = rel_err(rows, fx[name + "/rows"])
    worst["loss_mean"] = abs(float(mean) - float(fx[name + "/rows"].mean())) / max(1.0, abs(float(fx[name + "/rows"].mean())))
    for n in names:
        # the trainer's gradient smoothing on top of the reference's raw gradients
        # (packages/trainable/lua_src/supervised.lua:797-803)
        scale = 1.0 / math.sqrt(max(int(fx["%s/count/%s" % (name, n)]), 1) * bunch)
        want = fx["%s/g/%s" % (name, n)].astype(np.float64) * scale
        worst["grad " + n] = rel_err(np.asarray(tr.gradients(n)).reshape(want.shape), want)
    bad = {k: v for k, v in worst.items() if not v <= tol}
    assert not bad, (name, bad)
    return max(worst.values())
