#!/usr/bin/env python
"""Writes tests/golden/reference_steps.npz by RUNNING THE REFERENCE ITSELF (oracle/_ref/libaprilref.so, the
reference's own C++ compiled in place by oracle/ref_build/Makefile; needs /root/reference):

    python tests/golden/make_reference_fixtures.py

For every case of reference_cases.CASES: seeded input / target / weights, then -- from the reference's
ANN::*ANNComponent and ANN::*LossFunction objects -- the forward output, the per-row loss, the loss gradient,
the error back-propagated to the input, and the raw (unsmoothed) weight gradients with their shared counts.
The fixture travels to the GPU box, the reference does not.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import ref as R  # noqa: E402
from reference_cases import CASES, build_reference, weight_names  # noqa: E402

f32 = np.float32


def inputs_of(name, input_size, output_size, bunch, target_kind):
    rng = np.random.default_rng(sum(map(ord, name)))
    x = rng.uniform(0, 1, (bunch, input_size)).astype(f32)
    if target_kind == "onehot":
        t = np.eye(output_size, dtype=f32)[rng.integers(0, output_size, bunch)]
    elif target_kind == "binary":
        t = (rng.uniform(0, 1, (bunch, output_size)) > 0.5).astype(f32)
    else:
        t = rng.uniform(-1, 1, (bunch, output_size)).astype(f32)
    return rng, x, t


def main():
    out = {}
    for name, (layers, isz, osz, bunch, loss, tk) in CASES.items():
        rng, x, t = inputs_of(name, isz, osz, bunch, tk)
        net = build_reference(R, layers, isz, osz if not any(l[0] == "flatten" for l in layers[-1:]) else 0)
        net.forward(np.zeros((1, isz), f32), False)   # convolutions size their weights at the first forward
        net.reset(0)
        names = weight_names(layers)
        for n in names:
            shape = net.weight(n).shape
            fan = max(shape[1] if len(shape) > 1 and shape[1] > 1 else shape[0], 1)
            lo, hi = (0.05, 0.4) if n.startswith("a") else (-1.0 / np.sqrt(fan), 1.0 / np.sqrt(fan))
            w = rng.uniform(lo, hi, shape).astype(f32)
            net.set_weight(n, w)
            out["%s/w/%s" % (name, n)] = w
        y = net.forward(x, True)
        L = R.Loss(loss, osz)
        rows = L.loss_rows(y, t)
        g = L.gradient(y, t)
        dx = net.backprop(g)
        net.compute_gradients()
        out["%s/x" % name], out["%s/t" % name] = x, t
        out["%s/y" % name], out["%s/rows" % name], out["%s/lossgrad" % name] = y, rows, g
        out["%s/dx" % name] = dx.reshape(bunch, -1)
        for n in names:
            out["%s/g/%s" % (name, n)] = net.gradient(n)
            out["%s/count/%s" % (name, n)] = np.int32(net.shared_count(n))
        net.close()
        print("%-26s y%s loss mean %.6f" % (name, y.shape, float(rows.mean())))
    path = os.path.join(HERE, "reference_steps.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
