#!/usr/bin/env python
"""Runs the REFERENCE's own components on the GPU through the integration shims.

    APRILREF_LIB=integration/_build/libaprilref_b200.so python tests/ref_on_b200.py

libaprilref_b200.so is the reference's USE_CUDA build (integration/Makefile): its
Matrix / GPUMirroredMemoryBlock / component / loss code compiled by nvcc for
sm_100a, with mathcore's gemm / gemv / axpy translation units replaced by
integration/mathcore/*_b200.cc, i.e. dense fp32 BLAS goes through libb200ann.so.
Each network below is built twice from the reference's classes, once with
set_use_cuda(false) and once with set_use_cuda(true); forward outputs,
back-propagated errors and weight gradients must agree (fp32 math mode, 1e-4 of
the tensor's largest magnitude).  Prints one JSON line; exit status 0 = agreed.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("APRILREF_LIB", os.path.join(ROOT, "integration", "_build", "libaprilref_b200.so"))

from oracle import ref as R  # noqa: E402  (test infrastructure: the C face of the reference build)

f32 = np.float32
TOL = 1e-4


def err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.abs(a - b).max() / max(1.0, np.abs(b).max()))


def mlp(cuda, topology, names):
    s = R.stack()
    for (i, n, act), (wn, bn) in zip(topology, names):
        R.push(s, R.hyperplane(i, n, wn, bn), R.actf(act))
    R.set_use_cuda(s, cuda)
    return R.Net(s, topology[0][0], topology[-1][1])


def mlp_component_seam(cuda, topology, names):
    """The same MLP, the device build out of B200DotProductANNComponent (integration/ann/): its dense
    forward / backprop / gradients call b200_linear_{fwd,bwd_data,bwd_weight} directly."""
    s = R.stack()
    for (i, n, act), (wn, bn) in zip(topology, names):
        dp = R.b200_dot_product(i, n, wn) if cuda else R.dot_product(i, n, wn)
        R.push(s, dp, R.bias(n, bn), R.actf(act))
    R.set_use_cuda(s, cuda)
    return R.Net(s, topology[0][0], topology[-1][1])


def conv(cuda):
    s = R.stack()
    R.push(s, R.rewrap([1, 12, 12]), R.convolution([1, 3, 3], 4, "cw1"), R.convolution_bias(3, 4, "cb1"),
           R.actf("relu"), R.max_pooling([1, 2, 2]), R.flatten(), R.hyperplane(100, 5, "w", "b"),
           R.actf("log_softmax"))
    R.set_use_cuda(s, cuda)
    net = R.Net(s, 144, 5)
    net.forward(np.zeros((1, 144), f32), False)  # the convolution sizes its weights at the first forward
    net.reset(0)
    return net


def step(net, names, w, x, t):
    for n in names:
        net.set_weight(n, w[n])
    y = net.forward(x, True)
    g = (np.exp(y) - t).astype(f32)  # the MCCE gradient on log_softmax outputs
    dx = net.backprop(g)
    net.compute_gradients()
    return [y, dx] + [net.gradient(n) for n in names]


def compare(label, build, names, x, t, rng):
    """Worst relative error of the device build against the host build; a string if the case could not run
    (one failing case must not hide the others' numbers)."""
    try:
        cpu, gpu = build(False), build(True)
        w = {n: rng.uniform(-0.5, 0.5, cpu.weight(n).shape).astype(f32) for n in names}
        worst = max(err(a, b) for a, b in zip(step(gpu, names, w, x, t), step(cpu, names, w, x, t)))
    except Exception as e:  # noqa: BLE001  (R.ReferenceError_ carries the reference's own message)
        return label, "error: %s" % str(e)[:300]
    return label, worst


def main():
    if not R.built_with_cuda() and not os.environ.get("REF_ON_B200_DRY_RUN"):  # dry run: plumbing check on a CPU build
        print(json.dumps({"error": "library is not a USE_CUDA build", "lib": R.LIB_PATH}))
        return 2
    rng = np.random.default_rng(7)
    out = {}
    onehot = lambda rows, c: np.eye(c, dtype=f32)[rng.integers(0, c, rows)]  # noqa: E731

    topo = [(256, 256, "tanh"), (256, 128, "tanh"), (128, 10, "log_softmax")]
    names3 = [("w1", "b1"), ("w2", "b2"), ("w3", "b3")]
    flat = ["w1", "b1", "w2", "b2", "w3", "b3"]
    for bunch in (32, 1):  # bunch 1 takes the gemv / ger path of DotProductANNComponent
        k, v = compare("mlp_digits_bunch%d" % bunch, lambda c: mlp(c, topo, names3), flat,
                       rng.uniform(0, 1, (bunch, 256)).astype(f32), onehot(bunch, 10), rng)
        out[k] = v
    k, v = compare("mlp_component_seam_bunch32", lambda c: mlp_component_seam(c, topo, names3), flat,
                   rng.uniform(0, 1, (32, 256)).astype(f32), onehot(32, 10), rng)
    out[k] = v
    k, v = compare("conv_maxpool", conv, ["cw1", "cb1", "w", "b"],
                   rng.uniform(0, 1, (6, 144)).astype(f32), onehot(6, 5), rng)
    out[k] = v
    a = rng.uniform(-1, 1, (40, 70)).astype(f32)
    b = rng.uniform(-1, 1, (50, 70)).astype(f32)
    c = rng.uniform(-1, 1, (40, 50)).astype(f32)
    want = 0.5 * a.astype(np.float64) @ b.T + 2.0 * c
    R.set_use_cuda_default(True)
    try:
        out["matGemm_NT"] = err(R.gemm(0, 1, 0.5, a, b, 2.0, c), want)
    except Exception as e:  # noqa: BLE001
        out["matGemm_NT"] = "error: %s" % str(e)[:300]
    R.set_use_cuda_default(False)
    ok = all(isinstance(v, float) and v <= TOL for v in out.values())
    print(json.dumps({"ok": ok, "tolerance": TOL, "max_rel_err": out}))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
