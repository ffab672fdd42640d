"""Pins the CPU oracle against the reference's own golden vectors (CPU only)."""
import numpy as np
import pytest

from oracle import MTRand
from oracle import april as A
from oracle.digits import load_digits

# TEST/digitos/test.lua:15-27 (same table: packages/ann/optimizer/test/test-digits-sgd.lua:37-49)
GOLDEN_DIGITS = [
    (2.2762842, 2.0276833), (1.6794761, 1.2444804), (0.9245928, 0.6157830),
    (0.5167769, 0.3807266), (0.3109381, 0.3248250), (0.2184281, 0.2167415),
    (0.1626369, 0.1783843), (0.1271410, 0.1495624), (0.1077118, 0.1718368),
    (0.0960633, 0.1591717),
]


def test_mt19937_matches_published_generator():
    # numpy's legacy RandomState is the published MT19937 with init_genrand seeding
    mine = MTRand(1234).raw(3000)
    ref = np.random.RandomState(1234).randint(0, 2 ** 32, size=3000, dtype=np.uint64).astype(np.uint32)
    assert (mine == ref).all()
    # known answer: first output of MT19937 seeded with 5489 (the generator's default seed)
    assert MTRand(5489).randInt32() == 3499211612


def test_shuffle_is_permutation_and_deterministic():
    a = MTRand(5678).shuffle(800)
    assert sorted(a) == list(range(800))
    assert a == MTRand(5678).shuffle(800)


def build_digits_trainer(bunch=64):
    net = A.mlp_all_all("256 inputs 256 tanh 128 tanh 10 log_softmax")
    tr = A.SupervisedTrainer(net, A.MultiClassCrossEntropy(), bunch).build()
    tr.set_option("learning_rate", 0.08)
    tr.set_option("momentum", 0.01)
    tr.set_option("weight_decay", 1e-05)
    tr.set_layerwise_option("b.", "weight_decay", 0)
    tr.randomize_weights(random=MTRand(1234), inf=-1, sup=1, use_fanin=True)
    return tr


def test_digits_golden_curve():
    """TEST/digitos/test.lua: 10 epochs of (train, validation) MCCE loss, |d| <= 1e-3."""
    xtr, ttr, xva, tva = load_digits()
    tr = build_digits_trainer()
    shuffle = MTRand(5678)
    for epoch in range(10):
        trl, _ = tr.train_dataset(xtr, ttr, shuffle=shuffle)
        val, _ = tr.validate_dataset(xva, tva)
        assert abs(trl - GOLDEN_DIGITS[epoch][0]) <= 1e-3, (epoch, trl)
        assert abs(val - GOLDEN_DIGITS[epoch][1]) <= 1e-3, (epoch, val)


def build_conv_digits_trainer(bunch=64):
    """packages/ann/ann/test/test-convolution-digits.lua:69-137."""
    net = A.Stack()
    net.push(A.Rewrap((1, 16, 16)))
    net.push(A.Convolution((1, 3, 3), 10, "w1"))
    net.push(A.ConvolutionBias(10, "b1"))
    net.push(A.Actf("relu"))
    net.push(A.MaxPooling((1, 2, 2)))
    net.push(A.Convolution((10, 2, 2), 20, "w2"))
    net.push(A.ConvolutionBias(20, "b2"))
    net.push(A.Actf("relu"))
    net.push(A.MaxPooling((1, 2, 2)))
    net.push(A.Flatten())
    A.hyperplane(net, 180, 100, "w3", "b3")
    net.push(A.Actf("relu"))
    A.hyperplane(net, 100, 10, "w4", "b4")
    net.push(A.Actf("log_softmax"))
    tr = A.SupervisedTrainer(net, A.MultiClassCrossEntropy(), bunch).build(256)
    tr.set_option("learning_rate", 0.1)
    tr.set_option("momentum", 0.2)
    tr.set_option("weight_decay", 0.01)
    tr.set_option("L1_norm", 0.0)
    tr.set_option("max_norm_penalty", 4)
    for o in ("weight_decay", "max_norm_penalty", "L1_norm"):
        tr.set_layerwise_option("b.", o, 0.0)
    rnd = MTRand(1234)
    tr.randomize_weights(random=rnd, inf=-2.4, sup=2.4, use_fanin=True, use_fanout=True)
    tr.randomize_weights(name_match="b.", random=rnd, inf=0, sup=0.2, use_fanin=True, use_fanout=True)
    return tr


def test_conv_digits_initial_validation_loss():
    """First line of packages/ann/ann/test/test-convolution-digits-output.log:
    '# Initial validation error: 2.3320939540863' -- pins init order, convolution,
    convolution_bias, relu, max_pooling, flatten, hyperplane, log_softmax and MCCE
    forward.  (The per-epoch lines of that log are unasserted output of an older
    revision and are not reproduced by the current reference code either.)"""
    _, _, xva, tva = load_digits()
    tr = build_conv_digits_trainer()
    val, _ = tr.validate_dataset(xva, tva)
    assert abs(val - 2.3320939540863) < 2e-6


def _numeric_grad_check(tr, x, t, eps=1e-3, rel=0.1, max_checks=40):
    """The reference's own check: supervised.lua:907-991 (central differences,
    epsilon 1e-3, 10% relative tolerance), on the unsmoothed summed gradients."""
    tr.smooth_gradients = False
    tr.optimizer.set_option("learning_rate", 0.0)  # do not move the weights
    tr.optimizer.global_options["decay"] = 0.0
    tr.train_step(x, t)
    grads = {k: v.copy() for k, v in tr.grads.items()}
    rng = np.random.RandomState(0)
    bunch = x.shape[0]
    for name, w in tr.weights.items():
        flat = w.reshape(-1)
        for idx in rng.choice(flat.size, size=min(max_checks, flat.size), replace=False):
            orig = flat[idx]
            flat[idx] = orig - np.float32(eps)
            la = tr.loss.loss_rows(tr.net.forward(x, True), t).astype(np.float64).mean()
            flat[idx] = orig + np.float32(eps)
            lb = tr.loss.loss_rows(tr.net.forward(x, True), t).astype(np.float64).mean()
            flat[idx] = orig
            g = (lb - la) / (2 * eps)
            ann_g = float(grads[name].reshape(-1)[idx]) / bunch
            if abs(ann_g) > 2 * eps or abs(g) > 2 * eps:
                abs_err = abs(ann_g - g)
                err = 2 * abs_err / (abs(ann_g) + abs(g))
                assert not (err > rel and abs_err > 2 * eps), (name, idx, g, ann_g)


def test_conv_maxpool_gradients_numeric():
    """packages/ann/ann/test/test-components.lua:250-270 style check on the oracle's
    convolution / convolution_bias / max_pooling / flatten backward."""
    xtr, ttr, _, _ = load_digits()
    tr = build_conv_digits_trainer(8)
    _numeric_grad_check(tr, xtr[:8].copy(), ttr[:8].copy())


@pytest.mark.parametrize("actf", ["logistic", "tanh", "relu", "softmax"])
def test_mlp_gradients_numeric(actf):
    rnd = MTRand(7)
    topo = "6 inputs 5 %s 4 %s" % (actf, "softmax" if actf == "softmax" else "log_softmax")
    net = A.mlp_all_all(topo)
    loss = A.MSE() if actf == "softmax" else A.MultiClassCrossEntropy()
    tr = A.SupervisedTrainer(net, loss, 3).build()
    tr.randomize_weights(random=rnd, inf=-1, sup=1)
    x = (rnd.rand_array(18, 2.0) - 1.0).astype(np.float32).reshape(3, 6)
    t = np.zeros((3, 4), dtype=np.float32)
    t[np.arange(3), [1, 3, 0]] = 1
    _numeric_grad_check(tr, x, t)


def test_loss_closed_forms():
    """packages/ann/loss/test/test.lua:26-76: loss value and gradient against the
    formulas written in that test, on seeded inputs (random(1234) / random(525))."""
    r1, r2 = MTRand(1234), MTRand(525)
    # MSE
    i = r1.rand_array(80, 1.0).astype(np.float32).reshape(20, 4)
    t = np.array([r2.randInt(0, 1) for _ in range(80)], dtype=np.float32).reshape(20, 4)
    l = A.MSE()
    e, _ = l.compute_loss(i, t)
    assert abs(e - float(((i - t) ** 2).sum() * 0.5 / 20)) < 1e-5
    assert np.allclose(l.gradient(i, t), i - t)
    # multi-class cross entropy on normalised log-probabilities
    p = MTRand(1234).rand_array(80, 1.0).astype(np.float32).reshape(20, 4)
    p = p / p.sum(axis=1, keepdims=True)
    logp = np.log(p).astype(np.float32)
    cls = np.array([MTRand(525 + k).randInt(0, 3) for k in range(20)])
    t = np.zeros((20, 4), dtype=np.float32)
    t[np.arange(20), cls] = 1
    l = A.MultiClassCrossEntropy()
    e, _ = l.compute_loss(logp, t)
    assert abs(e - float(-(logp * t).sum() / 20)) < 1e-4
    assert np.allclose(l.gradient(logp, t), np.exp(logp) - t, atol=1e-6)
    # cross entropy on log-logistic outputs
    o = MTRand(1234).rand_array(20, 1.0).astype(np.float32).reshape(20, 1)
    o = np.clip(o, 1e-3, 1 - 1e-3)
    logo = np.log(o).astype(np.float32)
    t = np.array([MTRand(525 + k).randInt(0, 1) for k in range(20)], dtype=np.float32).reshape(20, 1)
    l = A.CrossEntropy()
    e, _ = l.compute_loss(logo, t)
    a = (logo * t).sum()
    b = ((1 - t) * np.log(1 - np.exp(logo))).sum()
    assert abs(e - float((-a - b) / 20)) < 1e-4
    assert np.allclose(l.gradient(logo, t), np.exp(logo) - t, atol=1e-6)


def test_gemm_exact_integer_cases():
    """packages/basics/matrix/test/test_gemm.lua:4-30: small integer matrices whose
    products are exact in fp32 -- the oracle's NT/NN/TN contractions must be exact."""
    a = np.arange(1, 7, dtype=np.float32).reshape(2, 3)
    b = np.arange(1, 13, dtype=np.float32).reshape(3, 4)
    assert (a @ b == np.array([[38, 44, 50, 56], [83, 98, 113, 128]], dtype=np.float32)).all()
    d = A.DotProduct(3, 4, "w")
    w = {}
    d.build(3, w)
    w["w"][...] = b.T
    y = d.forward(a)
    assert (y == a @ b).all()
    dy = np.ones((2, 4), dtype=np.float32)
    assert (d.backprop(dy) == dy @ b.T).all()
    g, c = {}, {}
    d.compute_gradients(g, c)
    assert (g["w"] == dy.T @ a).all() and c["w"] == 1


def test_tanh_is_antisym_logistic():
    """cmath_overloads.h:981-989: the reference's 'tanh' is 2/(1+e^-x)-1 = tanh(x/2)."""
    x = np.linspace(-4, 4, 33, dtype=np.float32)
    assert np.allclose(A.antisym_logistic(x), np.tanh(x / 2), atol=1e-6)
    assert not np.allclose(A.antisym_logistic(x), np.tanh(x), atol=1e-2)


# ---- the reference's optimizer golden curves -------------------------------
# packages/ann/optimizer/test/test-digits-{adagrad,rmsprop,adadelta,l1}.lua: same data, topology and
# shuffle seed as the SGD test; 10 epochs of (train, validation) loss, relative tolerance 1 % (5 % for L1).

GOLDEN_ADAGRAD = [(3.1195538, 2.0739560), (1.7988385, 1.3194453), (1.1200441, 0.8742127),
                  (0.6137433, 0.4673896), (0.3618005, 0.5647477), (0.2704565, 0.2707466),
                  (0.1515355, 0.4728075), (0.1356713, 0.1466638), (0.5097507, 0.1911529),
                  (0.0728004, 0.1801198)]
GOLDEN_RMSPROP = [(2.2983048, 2.3337903), (1.8389291, 1.2901400), (1.0636393, 0.7778704),
                  (0.5792997, 0.4687783), (0.3335530, 0.3594898), (0.2247733, 0.2889563),
                  (0.1688207, 0.2484314), (0.1469830, 0.2232731), (0.1109474, 0.2146144),
                  (0.1013973, 0.2133277)]
GOLDEN_ADADELTA = [(2.3053319, 2.3015521), (2.1570034, 1.8272703), (1.7826253, 1.9765174),
                   (1.4031528, 1.0635141), (1.2507044, 0.9257209), (0.7426600, 0.7414427),
                   (0.7360206, 0.5097144), (0.6020166, 0.4537252), (0.3055784, 0.3659463),
                   (0.2493757, 0.3132612)]
GOLDEN_L1 = [(2.2798486, 2.0456107), (1.7129538, 1.2954185), (0.9751059, 0.6590891),
             (0.5633602, 0.4138165), (0.3464335, 0.3450162), (0.2428290, 0.2518864),
             (0.1867137, 0.1979152), (0.1466725, 0.1708217), (0.1282059, 0.1904573),
             (0.1170338, 0.1766910)]


def _number_eq(a, b, eps):
    """utest.check.number_eq (packages/basics/utest/lua_src/utest.lua:126-132)."""
    return (a == 0 and b == 0) or abs(a - b) / abs(a + b) < eps


def _run_curve(tr, golden, eps, epochs=10):
    xtr, ttr, xva, tva = load_digits()
    shuffle = MTRand(5678)
    out = []
    for epoch in range(epochs):
        trl, _ = tr.train_dataset(xtr, ttr, shuffle=shuffle)
        val, _ = tr.validate_dataset(xva, tva)
        out.append((trl, val))
        assert _number_eq(trl, golden[epoch][0], eps), (epoch, trl, golden[epoch][0])
        assert _number_eq(val, golden[epoch][1], eps), (epoch, val, golden[epoch][1])
    return out


def _digits_trainer_with(optimizer, inf, sup):
    net = A.mlp_all_all("256 inputs 256 tanh 128 tanh 10 log_softmax")
    tr = A.SupervisedTrainer(net, A.MultiClassCrossEntropy(), 64, optimizer).build()
    return tr, dict(random=MTRand(1234), inf=inf, sup=sup, use_fanin=True)


def _adagrad_trainer(scale=None):
    tr, rw = _digits_trainer_with(A.Adagrad(), -0.1, 0.1)
    tr.set_option("weight_decay", 0.001)
    tr.set_option("learning_rate", 0.01)
    tr.set_layerwise_option("b.", "weight_decay", 0)
    tr.randomize_weights(**rw)
    if scale is not None:
        for w in tr.weights.values():
            w *= np.float32(scale)
    return tr


def test_digits_golden_curve_adagrad():
    """test-digits-adagrad.lua:20-150.

    With eps = 1e-6 in the denominator and lr = 0.01 the AdaGrad step is close to a sign step, and the
    trajectory amplifies float rounding: the golden curve itself is not monotone (0.136 -> 0.510 -> 0.073).
    The first epoch (12 updates from count 0, then a validation pass) reproduces the golden to 1e-3; from
    the third epoch on the curve depends on the BLAS summation order.  That is shown, not assumed: the second
    half of the test rescales the initial weights by 1 + 1e-6 and finds the oracle's own curve moving by more
    than the reference's 1 % tolerance.  So the pin is the epochs before the divergence."""
    got = _run_curve(_adagrad_trainer(), GOLDEN_ADAGRAD, 0.01, epochs=1)
    assert abs(got[0][0] - GOLDEN_ADAGRAD[0][0]) < 1e-3 * GOLDEN_ADAGRAD[0][0]
    assert abs(got[0][1] - GOLDEN_ADAGRAD[0][1]) < 1e-3 * GOLDEN_ADAGRAD[0][1]

    xtr, ttr, xva, tva = load_digits()
    curves = []
    for scale in (None, 1.0 + 1e-6):
        tr, shuffle, c = _adagrad_trainer(scale), MTRand(5678), []
        for epoch in range(4):
            trl, _ = tr.train_dataset(xtr, ttr, shuffle=shuffle)
            val, _ = tr.validate_dataset(xva, tva)
            c.append((trl, val))
        curves.append(c)
    # second epoch: training loss still within the reference's tolerance of the golden
    assert _number_eq(curves[0][1][0], GOLDEN_ADAGRAD[1][0], 0.01)
    # ... while a 1e-6 perturbation already moves the later validation losses by more than that tolerance
    assert any(not _number_eq(curves[0][e][1], curves[1][e][1], 0.01) for e in range(1, 4))


def test_digits_golden_curve_rmsprop():
    """test-digits-rmsprop.lua:20-152."""
    tr, rw = _digits_trainer_with(A.RMSProp(), -0.1, 0.1)
    tr.set_option("learning_rate", 0.001)
    tr.set_option("momentum", 0.4)
    tr.set_option("weight_decay", 0.001)
    tr.set_layerwise_option("b.", "weight_decay", 0)
    tr.randomize_weights(**rw)
    _run_curve(tr, GOLDEN_RMSPROP, 0.01)


def test_digits_golden_curve_adadelta():
    """test-digits-adadelta.lua:20-148."""
    tr, rw = _digits_trainer_with(A.Adadelta(), -0.1, 0.1)
    tr.set_option("weight_decay", 0.001)
    tr.set_layerwise_option("b.", "weight_decay", 0)
    tr.randomize_weights(**rw)
    _run_curve(tr, GOLDEN_ADADELTA, 0.01)


def test_digits_golden_curve_l1():
    """test-digits-l1.lua:6-135: SGD with L1 truncation (base_optimizer.lua:28-41)."""
    tr, rw = _digits_trainer_with(None, -1, 1)
    tr.set_option("learning_rate", 0.08)
    tr.set_option("momentum", 0.0)
    tr.set_option("L1_norm", 0.001)
    tr.set_layerwise_option("b.", "L1_norm", 0)
    tr.randomize_weights(**rw)
    _run_curve(tr, GOLDEN_L1, 0.05)


@pytest.mark.parametrize("make,steps,opts", [
    (lambda: A.SGD(), 200, {"learning_rate": 0.01, "momentum": 0.02}),
    (lambda: A.Adagrad(), 20000, {}),
    (lambda: A.RMSProp(), 20000, {}),
    (lambda: A.Adadelta(), 20000, {}),
], ids=["sgd", "adagrad", "rmsprop", "adadelta"])
def test_convex_minimum(make, steps, opts):
    """The *ConvexTest of each optimizer test: minimise 3x^2 - 2x + 10 from x = -100, expect x = 0.333
    (check.eq on a matrix compares with 1e-3... the reference's matrix:equals default epsilon is 5 %)."""
    opt = make()
    for k, v in opts.items():
        opt.set_option(k, v)
    w = {"x": np.array([[-100.0]], np.float32)}
    for _ in range(steps):
        opt.before_eval(w)
        g = {"x": (6 * w["x"] - 2).astype(np.float32)}
        opt.execute(w, g)
    assert abs(float(w["x"][0, 0]) - 0.333) <= 0.05 * 0.333, float(w["x"][0, 0])
