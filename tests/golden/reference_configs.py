"""The BASELINE.json workloads built from the reference's own classes (oracle/ref.py over
oracle/_ref/libaprilref.so), with the weight names the product and the oracle use for the same networks
(april_ann_b200/configs.py, oracle/configs.py)."""
import math

import numpy as np

NAMES = {
    "C2": ["w1", "b1", "w2", "b2", "w3", "b3"],
    "C2tanh": ["w1", "b1", "w2", "b2", "w3", "b3"],
    "C4": ["w1", "b1", "w2", "b2", "w3", "b3", "w4", "b4"],
    "C5": ["w1", "b1"],
}
SHAPES = {"C2": (1024, 784, 10), "C2tanh": (1024, 784, 10), "C4": (512, 784, 10), "C5": (4096, 512, 10000)}


def reference_net(R, name):
    s = R.stack()
    if name in ("C2", "C2tanh"):
        a = "relu" if name == "C2" else "tanh"
        R.push(s, R.hyperplane(784, 2048, "w1", "b1"), R.actf(a), R.hyperplane(2048, 2048, "w2", "b2"), R.actf(a),
               R.hyperplane(2048, 10, "w3", "b3"), R.actf("log_softmax"))
    elif name == "C4":
        R.push(s, R.rewrap([1, 28, 28]),
               R.convolution([1, 5, 5], 16, "w1"), R.convolution_bias(3, 16, "b1"), R.actf("relu"), R.max_pooling([1, 2, 2]),
               R.convolution([16, 5, 5], 32, "w2"), R.convolution_bias(3, 32, "b2"), R.actf("relu"), R.max_pooling([1, 2, 2]),
               R.flatten(), R.hyperplane(512, 256, "w3", "b3"), R.actf("relu"), R.hyperplane(256, 10, "w4", "b4"),
               R.actf("log_softmax"))
    elif name == "C5":
        R.push(s, R.hyperplane(512, 10000, "w1", "b1"), R.actf("log_softmax"))
    else:
        raise ValueError(name)
    bunch, nin, nout = SHAPES[name]
    net = R.Net(s, nin, nout)
    if name == "C4":
        net.forward(np.zeros((1, nin), np.float32), False)   # the convolutions size their weights at the first forward
        net.reset(0)
    return net


def reference_step(R, name, weights, x, t):
    """forward, per-row MCCE, backward and raw weight gradients out of the reference's classes.
    Returns (y, rows, {name: raw gradient}, {name: smoothing scale 1/sqrt(shared_count * bunch)})."""
    bunch, nin, nout = SHAPES[name]
    net = reference_net(R, name)
    for n in NAMES[name]:
        net.set_weight(n, weights[n])
    y = net.forward(x, True, out_elems=bunch * nout)
    loss = R.Loss("multi_class_cross_entropy", nout)
    rows = loss.loss_rows(y, t)
    net.backprop(loss.gradient(y, t))
    net.compute_gradients()
    grads = {n: net.gradient(n) for n in NAMES[name]}
    scale = {n: 1.0 / math.sqrt(max(net.shared_count(n), 1) * bunch) for n in NAMES[name]}   # supervised.lua:797-803
    net.close()
    return y, rows, grads, scale


def inputs(name, seed=2026):
    bunch, nin, nout = SHAPES[name]
    rs = np.random.RandomState(seed)
    x = rs.uniform(0.0 if name == "C4" else -1.0, 1.0, size=(bunch, nin)).astype(np.float32)
    t = np.zeros((bunch, nout), np.float32)
    t[np.arange(bunch), rs.randint(0, nout, size=bunch)] = 1.0
    return x, t
