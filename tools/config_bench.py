"""Step time of the other BASELINE.json configurations at full size on one B200 (not bench lines: the
bench line is C2; these show that the same path runs them and what it reaches).
    python tools/config_bench.py [C1 C3 C4 C5]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import april_ann_b200 as ann  # noqa: E402
from april_ann_b200._lib import lib, check  # noqa: E402

ctx = ann.get_context(0)
ctx.set_math_mode(ann.MATH_TF32)
rng = np.random.RandomState(3)


def onehot(n, c):
    t = np.zeros((n, c), np.float32)
    t[np.arange(n), rng.randint(0, c, n)] = 1
    return t


def conv_c4():
    c = ann.components
    net = c.stack(name="stack")
    net.push(c.rewrap(size=(1, 28, 28), name="rewrap"),
             c.convolution(kernel=(1, 5, 5), n=16, name="conv-w1", weights="w1"),
             c.convolution_bias(n=16, ndims=3, name="conv-b1", weights="b1"),
             c.actf.relu(name="actf-1"), c.max_pooling(kernel=(1, 2, 2), name="pool-1"),
             c.convolution(kernel=(16, 5, 5), n=32, name="conv-w2", weights="w2"),
             c.convolution_bias(n=32, ndims=3, name="conv-b2", weights="b2"),
             c.actf.relu(name="actf-2"), c.max_pooling(kernel=(1, 2, 2), name="pool-2"),
             c.flatten(name="flatten"),
             c.hyperplane(input=512, output=256, name="hyp-1", bias_name="b3", dot_product_name="w3",
                          bias_weights="b3", dot_product_weights="w3"),
             c.actf.relu(name="actf-3"),
             c.hyperplane(input=256, output=10, name="hyp-2", bias_name="b4", dot_product_name="w4",
                          bias_weights="b4", dot_product_weights="w4"),
             c.actf.log_softmax(name="actf-4"))
    return net


CONFIGS = {
    # name: (network, loss, bunch, input size, output size, algorithmic FLOPs per step or None)
    "C1": (lambda: ann.mlp.all_all.generate("256 inputs 256 tanh 128 tanh 10 log_softmax"), "mcce", 32, 256, 10, None),
    "C3": (lambda: ann.mlp.all_all.generate("4096 inputs " + " ".join(["4096 tanh"] * 8)), "mse", 8192, 4096, 4096, None),
    "C4": (conv_c4, "mcce", 512, 784, 10, None),
    "C5": (lambda: ann.mlp.all_all.generate("512 inputs 10000 log_softmax"), "mcce", 4096, 512, 10000, None),
}


def flops_mlp(sizes, bunch):
    p = sum(a * b for a, b in zip(sizes, sizes[1:]))
    return 2 * bunch * p * 3 - 2 * bunch * sizes[0] * sizes[1]


for name in (sys.argv[1:] or ["C1", "C3", "C4", "C5"]):
    mk, lossn, bunch, nin, nout, _ = CONFIGS[name]
    loss = ann.loss.multi_class_cross_entropy() if lossn == "mcce" else ann.loss.mse()
    tr = ann.trainable.supervised_trainer(mk(), loss, bunch, ctx=ctx)
    tr.build()
    tr.set_option("learning_rate", 0.01)
    tr.set_option("momentum", 0.9)
    tr.set_option("weight_decay", 1e-4)
    tr.randomize_weights(random=ann.random(1234), inf=-1, sup=1, use_fanin=True, use_fanout=True)
    x = rng.uniform(-1, 1, (bunch, nin)).astype(np.float32)
    t = onehot(bunch, nout) if lossn == "mcce" else rng.uniform(-1, 1, (bunch, nout)).astype(np.float32)
    tr.stage(x, t, bunch)
    for _ in range(4):
        tr.step_staged(bunch)
    ctx.sync()
    e0, e1 = C.c_void_p(), C.c_void_p()
    check(lib.b200_event_create(C.byref(e0)))
    check(lib.b200_event_create(C.byref(e1)))
    reps = 5 if name == "C3" else 30
    check(lib.b200_event_record(ctx.h, e0))
    for _ in range(reps):
        tr.step_staged(bunch)
    check(lib.b200_event_record(ctx.h, e1))
    ms = C.c_float()
    check(lib.b200_event_elapsed_ms(e0, e1, C.byref(ms)))
    us = ms.value / reps * 1e3
    fl = {"C1": flops_mlp([256, 256, 128, 10], 32), "C3": flops_mlp([4096] * 9, 8192), "C5": flops_mlp([512, 10000], 4096)}.get(name)
    mean, _ = tr.loss_get()
    print("%s: bunch %d, %.1f us per step, %.0f samples/s%s, running loss %.4f" % (
        name, bunch, us, bunch / us * 1e6, (", %.0f TFLOP/s (algorithmic)" % (fl / us / 1e6)) if fl else "", mean), flush=True)
    del tr
