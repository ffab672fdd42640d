"""Pins oracle/april.py against THE REFERENCE ITSELF, compiled here.

oracle/_ref/libaprilref.so is the reference's own C++ components / loss
functions / matrix code, compiled in place from /root/reference by
oracle/ref_build/Makefile (no reference source is copied into this
repository).  Every test below builds the same thing twice -- once out of the
reference's classes through oracle/ref.py, once out of the numpy restatement
-- feeds both the same seeded input, and compares forward outputs, back-
propagated errors, weight gradients, shared counts and per-row losses.

fp32 tolerance: the two sides differ only in summation order (the reference
sums a row with its BLAS / iterator order, numpy pairwise), so 2e-6 relative
to the largest magnitude of the compared tensor, stated per assert.

CPU only (no `gpu` marker).  When neither the prebuilt library nor
/root/reference is present (the GPU box), the module is skipped.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import april as A  # noqa: E402
from oracle import ref as R  # noqa: E402
from oracle.mtrand import MTRand  # noqa: E402

f32 = np.float32


def _ensure_built():
    if R.available():
        return True
    if not os.path.isdir("/root/reference/packages"):
        return False
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle", "ref_build"), "-j8"], check=True,
                   stdout=subprocess.DEVNULL)
    return R.available()


pytestmark = pytest.mark.skipif(not _ensure_built(),
                                reason="oracle/_ref/libaprilref.so not built and /root/reference absent")


def close(a, b, tol=2e-6):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    scale = max(1.0, float(np.abs(b).max()) if b.size else 1.0)
    err = float(np.abs(a - b).max()) if a.size else 0.0
    assert err <= tol * scale, "max abs err %.3g (scale %.3g)" % (err, scale)


def rand_weights(net, names, rng, lo=-0.5, hi=0.5):
    w = {}
    for n in names:
        w[n] = rng.uniform(lo, hi, net.weight(n).shape).astype(f32)
        net.set_weight(n, w[n])
    return w


# --------------------------------------------------------------------- actf

ACTFS = [("logistic", {}), ("tanh", {}), ("relu", {}), ("linear", {}), ("softplus", {}),
         ("softsign", {}), ("log_logistic", {}), ("leaky_relu", {"leak": 0.01}),
         ("leaky_relu", {"leak": 0.3}), ("hardtanh", {"inf": -1.0, "sup": 1.0}),
         ("hardtanh", {"inf": -0.25, "sup": 0.75}), ("softmax", {}), ("log_softmax", {})]


@pytest.mark.parametrize("kind,params", ACTFS, ids=[k + "".join("_%s" % v for v in p.values()) for k, p in ACTFS])
def test_activation_forward_backward(kind, params):
    rng = np.random.default_rng(11)
    x = rng.uniform(-4, 4, (9, 13)).astype(f32)
    # the saturating ends, exact zero and the hardtanh / relu corners
    x[0, :6] = [0.0, -0.0, 30.0, -30.0, 1.0, -1.0]
    x[1, :4] = [88.0, -88.0, 0.75, -0.25]
    if kind in ("softmax", "log_softmax"):
        x[1, :2] = [20.0, -20.0]
    dy = rng.uniform(-1, 1, x.shape).astype(f32)
    p = [params.get("leak", params.get("inf", 0.0)), params.get("sup", 0.0)]
    net = R.Net(R.push(R.stack(), R.actf(kind, *p)), 13, 13)
    o = A.Actf(kind, **params)
    close(net.forward(x, True), o.forward(x, True))
    close(net.backprop(dy), o.backprop(dy))
    net.close()


def test_softmax_wide_rows():
    # rows as wide as C5's, logits spread enough that most of the mass is in a few classes
    rng = np.random.default_rng(5)
    x = (rng.standard_normal((3, 10000)) * 6).astype(f32)
    for kind in ("softmax", "log_softmax"):
        net = R.Net(R.push(R.stack(), R.actf(kind)), 10000, 10000)
        close(net.forward(x, False), A.Actf(kind).forward(x, False))
        net.close()


def test_prelu():
    rng = np.random.default_rng(3)
    x = rng.uniform(-2, 2, (6, 10)).astype(f32)
    dy = rng.uniform(-1, 1, x.shape).astype(f32)
    for scalar in (False, True):
        net = R.Net(R.push(R.stack(), R.prelu(10, "a", scalar)), 10, 10)
        w = rand_weights(net, ["a"], rng, 0.05, 0.4)
        o = A.PReLU(10, "a", scalar)
        o.build(10, {"a": w["a"].copy()})
        close(net.forward(x, True), o.forward(x, True))
        close(net.backprop(dy), o.backprop(dy))
        net.compute_gradients()
        g, c = {}, {}
        o.compute_gradients(g, c)
        close(net.gradient("a"), g["a"])
        net.close()


def test_dropout_mask_stream():
    # same seed -> same MT19937 stream -> same mask, element by element
    x = np.random.default_rng(9).uniform(0.5, 1.5, (5, 40)).astype(f32)
    dy = np.ones_like(x)
    for prob, value in ((0.5, 0.0), (0.2, -1.0)):
        net = R.Net(R.push(R.stack(), R.dropout(4321, prob, value, True)), 40, 40)
        o = A.Dropout(MTRand(4321), prob, value, True)
        for _ in range(3):  # consecutive masks continue the same stream
            close(net.forward(x, True), o.forward(x, True), 0)
            close(net.backprop(dy), o.backprop(dy), 0)
            net.reset(0)
        close(net.forward(x, False), o.forward(x, False))
        net.close()


# --------------------------------------------------------------------- loss

def one_hot(rng, rows, classes):
    return np.eye(classes, dtype=f32)[rng.integers(0, classes, rows)]


def test_mse():
    rng = np.random.default_rng(1)
    o_, t_ = rng.uniform(-1, 1, (12, 7)).astype(f32), rng.uniform(-1, 1, (12, 7)).astype(f32)
    L, Lo = R.Loss("mse", 7), A.MSE()
    close(L.loss_rows(o_, t_), Lo.loss_rows(o_, t_))
    close(L.gradient(o_, t_), Lo.gradient(o_, t_))


def test_multi_class_cross_entropy_with_clamps():
    rng = np.random.default_rng(2)
    logits = (rng.standard_normal((16, 9)) * 8).astype(f32)  # some log-probabilities below ln(1e-6)
    logp = A.log_softmax_rows(logits)
    t_ = one_hot(rng, 16, 9)
    t_[3] = 0.0  # a row without a target
    t_[4] = [0.5, 0.5, 0, 0, 0, 0, 0, 0, 0]  # a soft target
    L, Lo = R.Loss("multi_class_cross_entropy", 9), A.MultiClassCrossEntropy()
    close(L.loss_rows(logp, t_), Lo.loss_rows(logp, t_))
    close(L.gradient(logp, t_), Lo.gradient(logp, t_))


def test_cross_entropy_on_log_logistic_outputs():
    rng = np.random.default_rng(4)
    z = (rng.standard_normal((10, 6)) * 5).astype(f32)
    logy = A.log_logistic(z)
    t_ = (rng.uniform(0, 1, z.shape) > 0.5).astype(f32)
    L, Lo = R.Loss("cross_entropy", 6), A.CrossEntropy()
    close(L.loss_rows(logy, t_), Lo.loss_rows(logy, t_), 5e-6)
    close(L.gradient(logy, t_), Lo.gradient(logy, t_))


def test_zero_one():
    rng = np.random.default_rng(6)
    # multi-class: target holds the 1-based class index in one column
    logp = A.log_softmax_rows(rng.standard_normal((20, 5)).astype(f32))
    tgt = rng.integers(1, 6, (20, 1)).astype(f32)
    close(R.Loss("zero_one", 5).loss_rows(logp, tgt), A.ZeroOne().loss_rows(logp, tgt), 0)
    # binary: one output against the threshold
    p = rng.uniform(0, 1, (20, 1)).astype(f32)
    p[0, 0] = 0.5
    tb = (rng.uniform(0, 1, (20, 1)) > 0.5).astype(f32)
    for th in (0.5, 0.3):
        close(R.Loss("zero_one", 1, th).loss_rows(p, tb), A.ZeroOne(th).loss_rows(p, tb), 0)


# ---------------------------------------------------------------- networks

def mlp_pair(topology, names, rng):
    """topology = [(in, out, actf), ...]; names = [(wname, bname), ...]."""
    s, o = R.stack(), A.Stack()
    for (i, n, act), (wn, bn) in zip(topology, names):
        R.push(s, R.hyperplane(i, n, wn, bn), R.actf(act))
        A.hyperplane(o, i, n, wn, bn)
        o.push(A.Actf(act))
    net = R.Net(s, topology[0][0], topology[-1][1])
    uniq = sorted({n for pair in names for n in pair})
    w = rand_weights(net, uniq, rng)
    ow = {k: v.copy() for k, v in w.items()}
    o.build(topology[0][0], ow)
    return net, o, uniq


def check_step(net, o, names, x, t_, loss_kind, oracle_loss, tol=2e-6):
    y, yo = net.forward(x, True), o.forward(x, True)
    close(y, yo, tol)
    L = R.Loss(loss_kind, y.shape[1])
    close(L.loss_rows(y, t_), oracle_loss.loss_rows(yo, t_), tol)
    g, go = L.gradient(y, t_), oracle_loss.gradient(yo, t_)
    close(g, go, tol)
    close(net.backprop(g), o.backprop(go), tol)
    net.compute_gradients()
    G, Cn = {}, {}
    o.compute_gradients(G, Cn)
    for n in names:
        close(net.gradient(n), G[n], tol)
        assert net.shared_count(n) == Cn[n], n


def test_mlp_step_digits_topology():
    # TEST/digitos/test.lua:5 -- 256 inputs, 256 tanh, 128 tanh, 10 log_softmax, bunch 32
    rng = np.random.default_rng(21)
    net, o, names = mlp_pair([(256, 256, "tanh"), (256, 128, "tanh"), (128, 10, "log_softmax")],
                             [("w1", "b1"), ("w2", "b2"), ("w3", "b3")], rng)
    x = rng.uniform(0, 1, (32, 256)).astype(f32)
    check_step(net, o, names, x, one_hot(rng, 32, 10), "multi_class_cross_entropy",
               A.MultiClassCrossEntropy(), 4e-6)
    net.close()


def test_mlp_relu_logistic_mse():
    rng = np.random.default_rng(22)
    net, o, names = mlp_pair([(20, 33, "relu"), (33, 17, "logistic"), (17, 4, "linear")],
                             [("w1", "b1"), ("w2", "b2"), ("w3", "b3")], rng)
    x = rng.uniform(-1, 1, (11, 20)).astype(f32)
    check_step(net, o, names, x, rng.uniform(-1, 1, (11, 4)).astype(f32), "mse", A.MSE())
    net.close()


def test_shared_weights_count_and_accumulate():
    # the same weight matrix in two layers: gradients add up, shared count is 2
    rng = np.random.default_rng(23)
    net, o, names = mlp_pair([(12, 12, "tanh"), (12, 12, "tanh"), (12, 3, "log_softmax")],
                             [("w", "b"), ("w", "b"), ("w3", "b3")], rng)
    x = rng.uniform(-1, 1, (8, 12)).astype(f32)
    check_step(net, o, names, x, one_hot(rng, 8, 3), "multi_class_cross_entropy", A.MultiClassCrossEntropy())
    assert net.shared_count("w") == 2
    net.close()


def test_bunch_of_one_takes_the_gemv_ger_path():
    # dot_product_component.cc:82,145,210 switch to gemv / ger when the bunch is 1
    rng = np.random.default_rng(24)
    net, o, names = mlp_pair([(9, 7, "tanh"), (7, 3, "log_softmax")], [("w1", "b1"), ("w2", "b2")], rng)
    x = rng.uniform(-1, 1, (1, 9)).astype(f32)
    check_step(net, o, names, x, one_hot(rng, 1, 3), "multi_class_cross_entropy", A.MultiClassCrossEntropy())
    net.close()


def conv_pair(rng, hw=12, k1=3, n1=4, k2=3, n2=6, pool=2, classes=5, step1=1):
    s, o = R.stack(), A.Stack()
    R.push(s, R.rewrap([1, hw, hw]),
           R.convolution([1, k1, k1], n1, "cw1", [1, step1, step1]), R.convolution_bias(3, n1, "cb1"), R.actf("relu"),
           R.max_pooling([1, pool, pool]),
           R.convolution([n1, k2, k2], n2, "cw2"), R.convolution_bias(3, n2, "cb2"), R.actf("tanh"),
           R.flatten())
    for c in (A.Rewrap((1, hw, hw)), A.Convolution((1, k1, k1), n1, "cw1", (1, step1, step1)),
              A.ConvolutionBias(n1, "cb1"), A.Actf("relu"), A.MaxPooling((1, pool, pool)),
              A.Convolution((n1, k2, k2), n2, "cw2"), A.ConvolutionBias(n2, "cb2"), A.Actf("tanh"), A.Flatten()):
        o.push(c)
    h1 = ((hw - k1) // step1 + 1) // pool
    h2 = h1 - k2 + 1
    flat = n2 * h2 * h2
    R.push(s, R.hyperplane(flat, classes, "w", "b"), R.actf("log_softmax"))
    A.hyperplane(o, flat, classes, "w", "b")
    o.push(A.Actf("log_softmax"))
    net = R.Net(s, hw * hw, classes)
    names = ["cw1", "cb1", "cw2", "cb2", "w", "b"]
    # the convolution weights only exist after the first forward has seen the image size
    net.forward(np.zeros((1, hw * hw), f32), False)
    net.reset(0)
    w = rand_weights(net, names, rng)
    o.build(hw * hw, {k: v.copy() for k, v in w.items()})
    return net, o, names


@pytest.mark.parametrize("step1", [1, 2])
def test_convolution_maxpool_net_step(step1):
    rng = np.random.default_rng(31 + step1)
    net, o, names = conv_pair(rng, hw=14 if step1 == 1 else 17, step1=step1)
    hw = 14 if step1 == 1 else 17
    x = rng.uniform(0, 1, (6, hw * hw)).astype(f32)
    check_step(net, o, names, x, one_hot(rng, 6, 5), "multi_class_cross_entropy",
               A.MultiClassCrossEntropy(), 4e-6)
    net.close()


def test_maxpool_ties_first_maximum_wins():
    # a constant image: every window is a tie, the error goes to the first element of each window
    s = R.push(R.stack(), R.rewrap([1, 4, 4]), R.max_pooling([1, 2, 2]), R.flatten())
    net = R.Net(s, 16, 0)  # a stack ending in flatten has no static output size
    o = A.Stack()
    for c in (A.Rewrap((1, 4, 4)), A.MaxPooling((1, 2, 2)), A.Flatten()):
        o.push(c)
    o.build(16, {})
    x = np.full((2, 16), 0.25, f32)
    e = np.arange(8, dtype=f32).reshape(2, 4) + 1
    close(net.forward(x, True), o.forward(x, True), 0)
    close(net.backprop(e), o.backprop(e), 0)
    net.close()


# --------------------------------------------------------------------- gemm

@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_gemm_seam(ta, tb):
    # AprilMath::doGemm through MatrixExt::BLAS::matGemm, the seam b200_sgemm replaces
    rng = np.random.default_rng(40 + 2 * ta + tb)
    m, n, k = 7, 5, 9
    a = rng.uniform(-1, 1, (k, m) if ta else (m, k)).astype(f32)
    b = rng.uniform(-1, 1, (n, k) if tb else (k, n)).astype(f32)
    c = rng.uniform(-1, 1, (m, n)).astype(f32)
    want = 0.7 * ((a.T if ta else a).astype(np.float64) @ (b.T if tb else b)) + 0.3 * c
    close(R.gemm(ta, tb, 0.7, a, b, 0.3, c), want)
    close(R.gemm(ta, tb, 1.0, a, b, 0.0, c), (a.T if ta else a).astype(np.float64) @ (b.T if tb else b))


# ------------------------------------------------------- BASELINE configs at full size

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from reference_configs import NAMES, SHAPES, inputs, reference_step  # noqa: E402


@pytest.mark.parametrize("name", ["C2", "C2tanh", "C4", "C5"])
def test_full_size_step(name):
    """The sizes bench.py's parity object checks the GPU at: the oracle used there agrees with the reference's own
    classes on the whole step -- forward, per-row losses, every raw gradient (seconds of host time per side)."""
    from oracle import configs as OC
    bunch, nin, nout = SHAPES[name]
    if name == "C2tanh":
        tr = A.SupervisedTrainer(A.mlp_all_all("784 inputs 2048 tanh 2048 tanh 10 log_softmax"),
                                 A.MultiClassCrossEntropy(), bunch).build()
    else:
        tr = OC.build_trainer(name, bunch)
    # weights of the bench's scale, uniform(+-1/sqrt(fan_in + fan_out)); numpy's generator here (the MT19937 stream
    # of randomize_weights is pinned in test_random_stream_cpu.py and costs seconds at this size in pure Python)
    rs = np.random.RandomState(99)
    for n in NAMES[name]:
        shape = tr.weights[n].shape
        a = 1.0 / np.sqrt(shape[0] + shape[1])
        tr.weights[n][...] = rs.uniform(-a, a, size=shape).astype(f32)
    weights = {n: tr.weights[n].copy() for n in NAMES[name]}
    x, t = inputs(name)
    y_ref, rows_ref, g_ref, _ = reference_step(R, name, weights, x, t)
    y = tr.net.forward(x, True)
    close(y, y_ref, 4e-6)
    L = A.MultiClassCrossEntropy()
    close(L.loss_rows(y, t), rows_ref, 4e-6)
    tr.net.backprop(L.gradient(y, t))
    G, Cn = {}, {}
    tr.net.compute_gradients(G, Cn)
    # tanh / linear networks: every gradient to summation-order accuracy.  ReLU / max-pool networks: a handful of
    # the 10^6 pre-activations lie within 1e-6 of zero (C4 with these weights: 9 + 5 + 3), and a unit the two
    # summation orders put on different sides of zero flips its gate -- ONE flip moves a convolution's weight
    # gradient by ~1e-3 of its norm while the forward pass, the losses and every tensor above the flipped gate
    # still agree to 1e-6.  That discontinuity is the reference's own (activation_function_kernels.cu ReLU
    # derivative on x > 0); it is why the GPU tests of ReLU nets compare gate-aware (tests/test_gpu_fullsize.py).
    grad_tol = 4e-6 if name in ("C2tanh", "C5") else 5e-3
    for n in NAMES[name]:
        close(G[n], g_ref[n], grad_tol)
    if name == "C4":   # above the last flipped gate the agreement is exact again
        for n in ("w3", "b3", "w4", "b4"):
            close(G[n], g_ref[n], 4e-6)
