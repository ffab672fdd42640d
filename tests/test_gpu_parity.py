"""GPU parity tests: every C-ABI entry point of the hot path against the CPU oracle on the same
seeded inputs, plus the reference's golden vectors end to end.  All calls go through
libb200ann.so (ctypes).  Tolerances:
  * fp32 (FFMA) mode: rel-L2 <= 1e-5, the reference's own golden tolerance 1e-3 for the curve;
  * TF32 (tcgen05) mode: rel-L2 <= 2e-3 per contraction, curve within 1 %.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import MTRand  # noqa: E402
from oracle import april as A  # noqa: E402
from oracle.digits import load_digits  # noqa: E402

F32_TOL = 1e-5
TF32_TOL = 2e-3


@pytest.fixture(scope="module")
def ann():
    import april_ann_b200 as ann
    ann.get_context().set_math_mode(ann.MATH_FP32)
    return ann


@pytest.fixture(scope="module")
def ops(ann):
    from april_ann_b200 import ops
    return ops


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def rnd_mat(seed, *shape, lo=-1.0, hi=1.0):
    r = MTRand(seed)
    n = int(np.prod(shape))
    return (r.rand_array(n, hi - lo) + lo).astype(np.float32).reshape(shape)


# ------------------------------------------------------------------ BLAS seam
def test_gemm_exact_integer_cases(ops):
    """packages/basics/matrix/test/test_gemm.lua:4-30: integer-valued products are exact."""
    a = np.arange(1, 7, dtype=np.float32).reshape(2, 3)
    b = np.arange(1, 13, dtype=np.float32).reshape(3, 4)
    want = np.array([[38, 44, 50, 56], [83, 98, 113, 128]], dtype=np.float32)
    assert np.array_equal(ops.sgemm(0, 0, 1.0, a, b), want)
    assert np.array_equal(ops.sgemm(1, 0, 1.0, np.ascontiguousarray(a.T), b), want)
    assert np.array_equal(ops.sgemm(0, 1, 1.0, a, np.ascontiguousarray(b.T)), want)
    assert np.array_equal(ops.sgemm(1, 1, 1.0, np.ascontiguousarray(a.T), np.ascontiguousarray(b.T)), want)
    c0 = np.ones((2, 4), dtype=np.float32)
    assert np.array_equal(ops.sgemm(0, 0, 2.0, a, b, beta=3.0, Cin=c0), 2 * want + 3)


@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (5, 7, 3), (130, 129, 17), (257, 64, 300), (64, 10, 2048)])
def test_gemm_all_layouts(ops, ta, tb, M, N, K):
    A_ = rnd_mat(1, *((K, M) if ta else (M, K)))
    B_ = rnd_mat(2, *((N, K) if tb else (K, N)))
    want = (A_.T if ta else A_).astype(np.float64) @ (B_.T if tb else B_).astype(np.float64)
    got = ops.sgemm(ta, tb, 1.0, A_, B_)
    assert rel_l2(got, want) < F32_TOL


# ------------------------------------------------------------------ fused dense layer
@pytest.mark.parametrize("act", [None, "logistic", "tanh", "relu"])
@pytest.mark.parametrize("M,N,K", [(32, 128, 256), (7, 10, 33), (200, 300, 100)])
def test_linear_fwd(ops, act, M, N, K):
    X, W, b = rnd_mat(3, M, K), rnd_mat(4, N, K, lo=-0.2, hi=0.2), rnd_mat(5, N)
    z = (X @ W.T + b).astype(np.float32)
    want = {None: z, "logistic": A.logistic(z), "tanh": A.antisym_logistic(z), "relu": A.relu(z)}[act]
    assert rel_l2(ops.linear_fwd(X, W, b, act), want) < F32_TOL


@pytest.mark.parametrize("act", [None, "logistic", "tanh", "relu"])
def test_linear_bwd_data_fuses_previous_derivative(ops, act):
    M, N, K = 48, 70, 90
    dY, W = rnd_mat(6, M, N), rnd_mat(7, N, K, lo=-0.3, hi=0.3)
    yprev = {None: None, "logistic": A.logistic(rnd_mat(8, M, K, lo=-3, hi=3)),
             "tanh": A.antisym_logistic(rnd_mat(8, M, K, lo=-3, hi=3)), "relu": A.relu(rnd_mat(8, M, K))}[act]
    dx = (dY @ W).astype(np.float32)
    if act == "logistic":
        dx = A.logistic_der(yprev) * dx
    elif act == "tanh":
        dx = A.antisym_logistic_der(yprev) * dx
    elif act == "relu":
        dx = A.relu_der(yprev) * dx
    assert rel_l2(ops.linear_bwd_data(dY, W, act, yprev), dx) < F32_TOL


def test_linear_bwd_weight_scale_beta_and_bias(ops):
    M, N, K = 64, 33, 50
    dY, X = rnd_mat(9, M, N), rnd_mat(10, M, K)
    scale = float(1.0 / np.sqrt(M))
    dW, db = ops.linear_bwd_weight(dY, X, scale=scale)
    assert rel_l2(dW, scale * (dY.T @ X)) < F32_TOL
    assert rel_l2(db, scale * dY.sum(axis=0)) < F32_TOL
    dW0, db0 = rnd_mat(11, N, K), rnd_mat(12, N)
    dW2, db2 = ops.linear_bwd_weight(dY, X, scale=scale, beta=1.0, dW0=dW0, db0=db0)
    assert rel_l2(dW2, dW0 + scale * (dY.T @ X)) < F32_TOL
    assert rel_l2(db2, db0 + scale * dY.sum(axis=0)) < F32_TOL


@pytest.mark.parametrize("M,N,K", [(1024, 10, 2048), (37, 3, 50), (130, 16, 132), (1, 10, 64), (4096, 10, 512)])
@pytest.mark.parametrize("act", [None, "tanh", "relu"])
def test_skinny_output_layer_passes(ops, M, N, K, act):
    """Layers with <= 16 output neurons (the 10-class output layer of every BASELINE MLP) take the
    HBM-bound exact-fp32 kernels of skinny.cu: forward, data gradient with the previous layer's
    derivative, weight gradient with scale/beta and the fused bias gradient."""
    X, W, b = rnd_mat(90, M, K), rnd_mat(91, N, K, lo=-0.2, hi=0.2), rnd_mat(92, N)
    f = {None: lambda z: z, "tanh": A.antisym_logistic, "relu": A.relu}[act]
    want = f((X.astype(np.float64) @ W.T.astype(np.float64) + b).astype(np.float32))
    assert rel_l2(ops.linear_fwd(X, W, b, act), want) < F32_TOL
    dY = rnd_mat(93, M, N)
    yprev = f(rnd_mat(94, M, K, lo=-2, hi=2)) if act else None
    dx = (dY.astype(np.float64) @ W.astype(np.float64)).astype(np.float32)
    if act == "tanh":
        dx = A.antisym_logistic_der(yprev) * dx
    elif act == "relu":
        dx = A.relu_der(yprev) * dx
    assert rel_l2(ops.linear_bwd_data(dY, W, act, yprev), dx) < F32_TOL
    scale = float(1.0 / np.sqrt(M))
    dW0, db0 = rnd_mat(95, N, K), rnd_mat(96, N)
    dW, db = ops.linear_bwd_weight(dY, X, scale=scale, beta=1.0, dW0=dW0, db0=db0)
    assert rel_l2(dW, dW0 + scale * (dY.T.astype(np.float64) @ X.astype(np.float64))) < F32_TOL
    assert rel_l2(db, db0 + scale * dY.astype(np.float64).sum(axis=0)) < F32_TOL
    dW, db = ops.linear_bwd_weight(dY, X, scale=scale)
    assert rel_l2(dW, scale * (dY.T.astype(np.float64) @ X.astype(np.float64))) < F32_TOL
    assert rel_l2(db, scale * dY.astype(np.float64).sum(axis=0)) < F32_TOL


@pytest.mark.parametrize("M,N", [(1024, 2048), (1000, 33), (5, 4), (8192, 4096), (77, 130)])
def test_bias_gradient_column_sums(ops, M, N):
    """bias_component.cc:87-122 as a two-stage deterministic column sum."""
    dY = rnd_mat(97, M, N)
    X = rnd_mat(98, M, 8)
    _, db = ops.linear_bwd_weight(dY, X, scale=0.25)
    assert rel_l2(db, 0.25 * dY.astype(np.float64).sum(axis=0)) < F32_TOL
    db0 = rnd_mat(99, N)
    _, db = ops.linear_bwd_weight(dY, X, scale=0.25, beta=1.0, dW0=np.zeros((N, 8), np.float32), db0=db0)
    assert rel_l2(db, db0 + 0.25 * dY.astype(np.float64).sum(axis=0)) < F32_TOL


# ------------------------------------------------------------------ activations / row kernels
@pytest.mark.parametrize("act", ["logistic", "tanh", "relu"])
def test_actf_elementwise(ops, act):
    x = rnd_mat(13, 37, 101, lo=-6, hi=6)
    f = {"logistic": A.logistic, "tanh": A.antisym_logistic, "relu": A.relu}[act]
    y = ops.actf_fwd(act, x)
    assert np.allclose(y, f(x), atol=2e-7, rtol=1e-6)
    dy = rnd_mat(14, 37, 101)
    d = {"logistic": A.logistic_der, "tanh": A.antisym_logistic_der, "relu": A.relu_der}[act]
    assert np.allclose(ops.actf_bwd(act, y, dy), d(f(x)) * dy if act != "relu" else A.relu_der(x) * dy,
                       atol=2e-7, rtol=1e-6)


@pytest.mark.parametrize("C", [3, 10, 33, 300, 1500, 4097, 10000])
def test_softmax_and_log_softmax_rows(ops, C):
    x = rnd_mat(15, 19, C, lo=-8, hi=8)
    x[0, :] = 0.0            # flat row
    x[1, 0] = 60.0           # range > 30: the reference clamps the subtracted minimum
    assert rel_l2(ops.actf_fwd("softmax", x), A.softmax_rows(x)) < F32_TOL
    ls = ops.actf_fwd("log_softmax", x)
    assert np.allclose(ls, A.log_softmax_rows(x), atol=3e-6, rtol=1e-6)
    y = A.softmax_rows(x)
    dy = rnd_mat(16, 19, C)
    assert rel_l2(ops.actf_bwd("softmax", y, dy), A.softmax_der_rows(y, dy)) < 5e-5


@pytest.mark.parametrize("C", [4, 10, 1000])
def test_losses(ops, C):
    M = 23
    p = rnd_mat(17, M, C, lo=0.01, hi=1.0)
    p /= p.sum(axis=1, keepdims=True)
    logp = np.log(p).astype(np.float32)
    t = np.zeros((M, C), dtype=np.float32)
    t[np.arange(M), np.arange(M) % C] = 1.0
    rows, g = ops.loss_and_grad("multi_class_cross_entropy", logp, t)
    l = A.MultiClassCrossEntropy()
    assert np.allclose(rows, l.loss_rows(logp, t), rtol=2e-6, atol=1e-6)
    assert np.allclose(g, l.gradient(logp, t), atol=2e-7)
    o, tt = rnd_mat(18, M, C), rnd_mat(19, M, C, lo=0, hi=1)
    rows, g = ops.loss_and_grad("mse", o, tt)
    assert np.allclose(rows, A.MSE().loss_rows(o, tt), rtol=3e-6)
    assert np.array_equal(g, (o - tt).astype(np.float32))
    lo = np.log(np.clip(rnd_mat(20, M, C, lo=0, hi=1), 1e-4, 1 - 1e-4)).astype(np.float32)
    tb = (rnd_mat(21, M, C, lo=0, hi=1) > 0.5).astype(np.float32)
    rows, g = ops.loss_and_grad("cross_entropy", lo, tb)
    assert np.allclose(rows, A.CrossEntropy().loss_rows(lo, tb), rtol=2e-5, atol=1e-5)
    assert np.allclose(g, A.CrossEntropy().gradient(lo, tb), atol=2e-7)


@pytest.mark.parametrize("C", [10, 257, 10000])
def test_fused_log_softmax_mcce(ops, C):
    M = 31
    z = rnd_mat(22, M, C, lo=-5, hi=5)
    t = np.zeros((M, C), dtype=np.float32)
    t[np.arange(M), (np.arange(M) * 7) % C] = 1.0
    logp, rows, grad = ops.log_softmax_mcce_fused(z, t)
    want_logp = A.log_softmax_rows(z)
    l = A.MultiClassCrossEntropy()
    assert np.allclose(logp, want_logp, atol=3e-6)
    assert np.allclose(rows, l.loss_rows(want_logp, t), rtol=3e-6, atol=2e-6)
    assert np.allclose(grad, l.gradient(want_logp, t), atol=3e-7)


@pytest.mark.parametrize("M,N,K,act", [(64, 10, 2048, "relu"), (37, 3, 512, "tanh"), (8, 16, 1028, "logistic"),
                                       (1, 10, 64, None)])
def test_output_layer_fused_launch(ops, M, N, K, act):
    """dot_product + bias + log_softmax + MCCE + gradient + data gradient of the layer below, one launch,
    against the oracle's component-by-component result (ragged row counts, every N padding class)."""
    h = rnd_mat(40, M, K)
    if act:
        h = {"relu": A.relu, "tanh": A.antisym_logistic, "logistic": A.logistic}[act](h)
    w, b = rnd_mat(41, N, K, lo=-0.1, hi=0.1), rnd_mat(42, N)
    t = np.zeros((M, N), dtype=np.float32)
    t[np.arange(M), (np.arange(M) * 5) % N] = 1.0
    logits, logp, rows, grad, dx = ops.output_layer_fused(h, w, b, t, act_prev=act)
    want_logits = (h.astype(np.float64) @ w.astype(np.float64).T + b).astype(np.float32)
    assert rel_l2(logits, want_logits) < F32_TOL
    want_logp = A.log_softmax_rows(want_logits)
    l = A.MultiClassCrossEntropy()
    assert np.allclose(logp, want_logp, atol=2e-5)
    assert np.allclose(rows, l.loss_rows(want_logp, t), rtol=2e-5, atol=2e-5)
    want_grad = l.gradient(want_logp, t)
    assert np.allclose(grad, want_grad, atol=2e-6)
    want_dx = want_grad.astype(np.float64) @ w.astype(np.float64)
    if act:
        want_dx = want_dx * {"relu": A.relu_der, "tanh": A.antisym_logistic_der, "logistic": A.logistic_der}[act](h)
    assert rel_l2(dx, want_dx.astype(np.float32)) < 2e-5


# ------------------------------------------------------------------ convolution / pooling
@pytest.mark.parametrize("B,C,H,W,n,kh,kw,sh,sw", [(3, 1, 16, 16, 10, 3, 3, 1, 1), (2, 10, 7, 7, 20, 2, 2, 1, 1),
                                                   (4, 3, 12, 11, 5, 5, 4, 2, 3), (2, 16, 12, 12, 32, 5, 5, 1, 1)])
def test_convolution_all_passes(ops, B, C, H, W, n, kh, kw, sh, sw):
    x = rnd_mat(23, B, C, H, W)
    w = rnd_mat(24, n, C * kh * kw, lo=-0.3, hi=0.3)
    bias = rnd_mat(25, n)
    conv = A.Convolution((C, kh, kw), n, "w", step=(1, sh, sw))
    weights = {}
    conv.build(0, weights)
    weights["w"][...] = w
    y_ref = conv.forward(x)
    y = ops.conv2d_fwd(x, w, (kh, kw), (sh, sw))
    assert y.shape == y_ref.shape and rel_l2(y, y_ref) < F32_TOL
    y_fused = ops.conv2d_fwd(x, w, (kh, kw), (sh, sw), bias=bias, act="relu")
    assert rel_l2(y_fused, A.relu(y_ref + bias[None, :, None, None])) < F32_TOL
    dy = rnd_mat(26, *y_ref.shape)
    assert rel_l2(ops.conv2d_bwd_data(dy, w, x.shape, (kh, kw), (sh, sw)), conv.backprop(dy)) < F32_TOL
    g, c = {}, {}
    conv.compute_gradients(g, c)
    dw, db = ops.conv2d_bwd_weight(dy, x, (kh, kw), (sh, sw), scale=0.5)
    assert rel_l2(dw, 0.5 * g["w"]) < F32_TOL
    assert rel_l2(db, 0.5 * dy.sum(axis=(0, 2, 3))) < F32_TOL


@pytest.mark.parametrize("kernel,step", [((2, 2), None), ((3, 3), (2, 2)), ((2, 3), (1, 2))])
def test_max_pooling(ops, kernel, step):
    x = rnd_mat(27, 3, 4, 9, 10)
    x[0, 0, :4, :4] = 1.5  # ties: the first maximum must win
    mp = A.MaxPooling((1,) + kernel, (1,) + step if step else None)
    y_ref = mp.forward(x)
    y, arg = ops.maxpool_fwd(x, kernel, step)
    assert np.array_equal(y, y_ref)
    dy = rnd_mat(28, *y_ref.shape)
    assert np.allclose(ops.maxpool_bwd(dy, arg, x.shape, kernel, step), mp.backprop(dy), atol=1e-6)


# ------------------------------------------------------------------ trainer: step-level parity
def make_pair(ann, topo, loss_name, bunch, seed=1234, **opts):
    ref_loss = {"mcce": A.MultiClassCrossEntropy, "mse": A.MSE, "ce": A.CrossEntropy}[loss_name]()
    gpu_loss = {"mcce": ann.loss.multi_class_cross_entropy, "mse": ann.loss.mse, "ce": ann.loss.cross_entropy}[loss_name]()
    ref = A.SupervisedTrainer(A.mlp_all_all(topo), ref_loss, bunch).build()
    tr = ann.trainable.supervised_trainer(ann.mlp.all_all.generate(topo), gpu_loss, bunch).build()
    for k, v in opts.items():
        ref.set_option(k, v)
        tr.set_option(k, v)
    ref.set_layerwise_option("b.", "weight_decay", 0)
    tr.set_layerwise_option("b.", "weight_decay", 0)
    ref.randomize_weights(random=MTRand(seed), inf=-1, sup=1, use_fanin=True, use_fanout=True)
    tr.randomize_weights(random=ann.random(seed), inf=-1, sup=1, use_fanin=True, use_fanout=True)
    return ref, tr


def onehot(seed, M, C):
    r = MTRand(seed)
    t = np.zeros((M, C), dtype=np.float32)
    t[np.arange(M), [r.randInt(0, C - 1) for _ in range(M)]] = 1.0
    return t


@pytest.mark.parametrize("topo,loss_name", [
    ("20 inputs 16 tanh 12 logistic 5 log_softmax", "mcce"),
    ("20 inputs 24 relu 24 relu 5 log_softmax", "mcce"),
    ("20 inputs 16 logistic 5 softmax", "mse"),
    ("20 inputs 16 tanh 5 linear", "mse"),
])
@pytest.mark.parametrize("fuse,graph", [(1, 1), (1, 0), (0, 0)])
def test_train_steps_match_oracle(ann, topo, loss_name, fuse, graph):
    ref, tr = make_pair(ann, topo, loss_name, 9, learning_rate=0.05, momentum=0.6, weight_decay=1e-3)
    tr.set_flag("fuse", fuse)
    tr.set_flag("cuda_graph", graph)
    tr.set_flag("keep_gradients", 1)
    for n in tr.weight_names():
        assert np.array_equal(tr.weights(n), ref.weights[n]), n
    x, t = rnd_mat(30, 9, 20), onehot(31, 9, 5)
    for step in range(5):
        l_gpu, rows_gpu = tr.train_step(x, t)
        l_ref, rows_ref = ref.train_step(x, t)
        assert abs(l_gpu - l_ref) <= 2e-6 * max(1, abs(l_ref)), (step, l_gpu, l_ref)
        assert np.allclose(rows_gpu, rows_ref, rtol=5e-6, atol=2e-6)
        for n in tr.weight_names():
            assert rel_l2(tr.gradients(n), ref.grads[n]) < 2e-5, (step, n)
            assert rel_l2(tr.weights(n), ref.weights[n]) < F32_TOL, (step, n)
            assert rel_l2(tr.updates(n), ref.optimizer.update[n]) < 2e-5, (step, n)


def test_bunch_of_one_and_ragged_bunches(ann):
    """bunch_size == 1 takes the gemv/ger route in the reference (dot_product_component.cc:80-88)."""
    ref, tr = make_pair(ann, "12 inputs 8 tanh 4 log_softmax", "mcce", 4, learning_rate=0.1)
    for bunch in (1, 4, 3, 1, 2):
        x, t = rnd_mat(40 + bunch, bunch, 12), onehot(50 + bunch, bunch, 4)
        l_gpu, _ = tr.train_step(x, t)
        l_ref, _ = ref.train_step(x, t)
        assert abs(l_gpu - l_ref) < 2e-6 * max(1, abs(l_ref))
    for n in tr.weight_names():
        assert rel_l2(tr.weights(n), ref.weights[n]) < F32_TOL


def test_l1_and_max_norm_options(ann):
    ref, tr = make_pair(ann, "10 inputs 8 tanh 4 log_softmax", "mcce", 6, learning_rate=0.2, momentum=0.3,
                        L1_norm=1e-3, max_norm_penalty=0.9)
    x, t = rnd_mat(60, 6, 10), onehot(61, 6, 4)
    for _ in range(4):
        tr.train_step(x, t)
        ref.train_step(x, t)
    for n in tr.weight_names():
        assert rel_l2(tr.weights(n), ref.weights[n]) < 2e-5, n


def test_error_behaviour(ann):
    _, tr = make_pair(ann, "12 inputs 8 tanh 4 log_softmax", "mcce", 4)
    with pytest.raises(ann.B200Error) as e:
        tr.train_step(np.zeros((4, 11), np.float32), np.zeros((4, 4), np.float32))
    assert e.value.code == 128
    with pytest.raises(ann.B200Error):
        tr.set_option("no_such_option", 1.0)
    with pytest.raises(ann.B200Error):
        ann.mlp.all_all.generate("12 inputs 8 no_such_actf")
    with pytest.raises(ann.B200Error):
        ann.loss.multi_class_cross_entropy(2)
    unbuilt = ann.trainable.supervised_trainer(ann.mlp.all_all.generate("4 inputs 3 log_softmax"),
                                               ann.loss.multi_class_cross_entropy(), 2)
    with pytest.raises(ann.B200Error):
        unbuilt.train_step(np.zeros((2, 4), np.float32), np.zeros((2, 3), np.float32))


# ------------------------------------------------------------------ golden vectors end to end
GOLDEN_DIGITS = [
    (2.2762842, 2.0276833), (1.6794761, 1.2444804), (0.9245928, 0.6157830), (0.5167769, 0.3807266),
    (0.3109381, 0.3248250), (0.2184281, 0.2167415), (0.1626369, 0.1783843), (0.1271410, 0.1495624),
    (0.1077118, 0.1718368), (0.0960633, 0.1591717),
]


def digits_trainer(ann):
    tr = ann.trainable.supervised_trainer(ann.mlp.all_all.generate("256 inputs 256 tanh 128 tanh 10 log_softmax"),
                                          ann.loss.multi_class_cross_entropy(), 64).build()
    tr.set_option("learning_rate", 0.08)
    tr.set_option("momentum", 0.01)
    tr.set_option("weight_decay", 1e-05)
    tr.set_layerwise_option("b.", "weight_decay", 0)
    tr.randomize_weights(random=ann.random(1234), inf=-1, sup=1, use_fanin=True)
    return tr


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-3), ("tf32", 1e-2)])
def test_digits_golden_curve(ann, mode, tol):
    """TEST/digitos/test.lua:15-27,127-136 run verbatim through the B200 path: 10 epochs of
    (train, validation) loss within the reference's own epsilon (1e-3) in fp32 mode and within
    1 % (test-digits-sgd.lua:49's relative tolerance) in TF32 mode."""
    ann.get_context().set_math_mode(ann.MATH_TF32 if mode == "tf32" else ann.MATH_FP32)
    try:
        xtr, ttr, xva, tva = load_digits()
        tr = digits_trainer(ann)
        shuffle = ann.random(5678)
        for epoch in range(10):
            trl, _ = tr.train_dataset(xtr, ttr, shuffle=shuffle)
            val, _ = tr.validate_dataset(xva, tva)
            g = GOLDEN_DIGITS[epoch]
            assert abs(trl - g[0]) <= tol * (1 if mode == "fp32" else max(1.0, g[0] * 3)), (epoch, trl, g)
            assert abs(val - g[1]) <= tol * (1 if mode == "fp32" else max(1.0, g[1] * 3)), (epoch, val, g)
    finally:
        ann.get_context().set_math_mode(ann.MATH_FP32)


def conv_net(ann):
    c = ann.components
    net = c.stack(name="stack")
    net.push(c.rewrap(size=(1, 16, 16), name="rewrap"),
             c.convolution(kernel=(1, 3, 3), n=10, name="conv-w1", weights="w1"),
             c.convolution_bias(n=10, ndims=3, name="conv-b1", weights="b1"),
             c.actf.relu(name="actf-1"), c.max_pooling(kernel=(1, 2, 2), name="pool-1"),
             c.convolution(kernel=(10, 2, 2), n=20, name="conv-w2", weights="w2"),
             c.convolution_bias(n=20, ndims=3, name="conv-b2", weights="b2"),
             c.actf.relu(name="actf-2"), c.max_pooling(kernel=(1, 2, 2), name="pool-2"),
             c.flatten(name="flatten"),
             c.hyperplane(input=180, output=100, name="hyp-1", bias_name="b3", dot_product_name="w3",
                          bias_weights="b3", dot_product_weights="w3"),
             c.actf.relu(name="actf-3"),
             c.hyperplane(input=100, output=10, name="hyp-2", bias_name="b4", dot_product_name="w4",
                          bias_weights="b4", dot_product_weights="w4"),
             c.actf.log_softmax(name="actf-4"))
    return net


def test_conv_digits_against_golden_and_oracle(ann):
    """packages/ann/ann/test/test-convolution-digits.lua: initial validation loss of the log
    (2.3320939540863) and three epochs step for step against the oracle."""
    from test_oracle_golden import build_conv_digits_trainer
    xtr, ttr, xva, tva = load_digits()
    tr = ann.trainable.supervised_trainer(conv_net(ann), ann.loss.multi_class_cross_entropy(10), 64).build(256, 10)
    ref = build_conv_digits_trainer()
    for o, v in (("learning_rate", 0.1), ("momentum", 0.2), ("weight_decay", 0.01), ("L1_norm", 0.0),
                 ("max_norm_penalty", 4)):
        tr.set_option(o, v)
    for o in ("weight_decay", "max_norm_penalty", "L1_norm"):
        tr.set_layerwise_option("b.", o, 0.0)
    rnd = ann.random(1234)
    tr.randomize_weights(random=rnd, inf=-2.4, sup=2.4, use_fanin=True, use_fanout=True)
    tr.randomize_weights(name_match="b.", random=rnd, inf=0, sup=0.2, use_fanin=True, use_fanout=True)
    for n in tr.weight_names():
        assert np.array_equal(tr.weights(n), ref.weights[n]), n
    val, _ = tr.validate_dataset(xva, tva)
    assert abs(val - 2.3320939540863) < 5e-6
    sh_gpu, sh_ref = ann.random(5678), MTRand(5678)
    for epoch in range(3):
        trl, _ = tr.train_dataset(xtr, ttr, shuffle=sh_gpu)
        rl, _ = ref.train_dataset(xtr, ttr, shuffle=sh_ref)
        # this net trains chaotically (the reference's own log jumps 0.32 -> 1.09 -> 0.32 between
        # epochs 6-8): fp32 rounding differences grow with every step, so the epoch means are
        # compared loosely here and the strict comparison is the step-level test below
        tol = (2e-3, 2e-2, 1e-1)[epoch]
        assert abs(trl - rl) < tol * max(1.0, rl), (epoch, trl, rl)
        val, _ = tr.validate_dataset(xva, tva)
        rv, _ = ref.validate_dataset(xva, tva)
        assert abs(val - rv) < tol * max(1.0, rv), (epoch, val, rv)


def test_conv_net_single_step_matches_oracle(ann):
    from test_oracle_golden import build_conv_digits_trainer
    xtr, ttr, _, _ = load_digits()
    tr = ann.trainable.supervised_trainer(conv_net(ann), ann.loss.multi_class_cross_entropy(10), 16).build(256, 10)
    tr.set_flag("keep_gradients", 1)
    ref = build_conv_digits_trainer(16)
    for o, v in (("learning_rate", 0.1), ("momentum", 0.2), ("weight_decay", 0.01), ("max_norm_penalty", 4)):
        tr.set_option(o, v)
    for o in ("weight_decay", "max_norm_penalty", "L1_norm"):
        tr.set_layerwise_option("b.", o, 0.0)
    for n in tr.weight_names():
        tr.set_weights(n, ref.weights[n])
    for step in range(3):
        x, t = xtr[step * 16:(step + 1) * 16], ttr[step * 16:(step + 1) * 16]
        l_gpu, _ = tr.train_step(x, t)
        l_ref, _ = ref.train_step(x, t)
        assert abs(l_gpu - l_ref) < 5e-6 * max(1, abs(l_ref))
        for n in tr.weight_names():
            assert rel_l2(tr.gradients(n), ref.grads[n]) < 5e-5, (step, n)
            assert rel_l2(tr.weights(n), ref.weights[n]) < 2e-5, (step, n)


# ------------------------------------------------------------------ full-size properties
def test_c2_full_size_properties(ann):
    """BASELINE config #2 (784-2048-2048-10 relu/log_softmax, bunch 1024) at full size, checked
    through size-independent properties: the loss of the first step equals ln(10) within init
    noise, repeated steps on one bunch decrease the loss monotonically, weights stay finite and
    the two math modes agree within the TF32 tolerance."""
    losses = {}
    for mode in (ann.MATH_FP32, ann.MATH_TF32):
        ann.get_context().set_math_mode(mode)
        try:
            tr = ann.trainable.supervised_trainer(
                ann.mlp.all_all.generate("784 inputs 2048 relu 2048 relu 10 log_softmax"),
                ann.loss.multi_class_cross_entropy(), 1024).build()
            tr.set_option("learning_rate", 0.01)
            tr.set_option("momentum", 0.9)
            tr.set_option("weight_decay", 1e-4)
            tr.set_layerwise_option("b.", "weight_decay", 0)
            tr.randomize_weights(random=ann.random(1234), inf=-1, sup=1, use_fanin=True, use_fanout=True)
            x, t = rnd_mat(70, 1024, 784), onehot(71, 1024, 10)
            ls = [tr.train_step(x, t)[0] for _ in range(6)]
            assert abs(ls[0] - np.log(10)) < 0.2
            assert all(b < a for a, b in zip(ls, ls[1:])), ls
            for n in tr.weight_names():
                assert np.isfinite(tr.weights(n)).all()
            losses[mode] = ls
        finally:
            ann.get_context().set_math_mode(ann.MATH_FP32)
    assert np.allclose(losses[ann.MATH_FP32], losses[ann.MATH_TF32], rtol=5e-3)


# ------------------------------------------------------------------ tcgen05 TF32 contraction
@pytest.mark.parametrize("ta,tb", [(0, 1), (0, 0), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(128, 16, 32), (1000, 300, 100), (1024, 2048, 784), (2048, 784, 1024),
                                   (4096, 4096, 512)])
def test_tf32_tensor_core_gemm(ann, ops, ta, tb, M, N, K):
    """kind::tf32 tcgen05 path: exact on integer-valued operands (every product is exact in TF32,
    so a mismatch is a layout bug), within the stated TF32 tolerance on random operands, and
    measurably different from the fp32 FFMA result (proves the tensor path actually ran)."""
    ctx = ann.get_context()
    rng = np.random.RandomState(M + 3 * N + 7 * K + ta + 2 * tb)
    Ai = rng.randint(-3, 4, size=(K, M) if ta else (M, K)).astype(np.float32)
    Bi = rng.randint(-3, 4, size=(N, K) if tb else (K, N)).astype(np.float32)
    Ar = rng.uniform(-1, 1, size=Ai.shape).astype(np.float32)
    Br = rng.uniform(-1, 1, size=Bi.shape).astype(np.float32)
    want_i = (Ai.T if ta else Ai).astype(np.float64) @ (Bi.T if tb else Bi).astype(np.float64)
    want_r = (Ar.T if ta else Ar).astype(np.float64) @ (Br.T if tb else Br).astype(np.float64)
    ctx.set_math_mode(ann.MATH_TF32)
    try:
        got_i = ops.sgemm(ta, tb, 1.0, Ai, Bi)
        got_r = ops.sgemm(ta, tb, 1.0, Ar, Br)
    finally:
        ctx.set_math_mode(ann.MATH_FP32)
    assert np.array_equal(got_i, want_i.astype(np.float32))
    e_tf32 = rel_l2(got_r, want_r)
    assert e_tf32 < TF32_TOL
    e_fp32 = rel_l2(ops.sgemm(ta, tb, 1.0, Ar, Br), want_r)
    assert e_fp32 < F32_TOL and e_tf32 > 10 * e_fp32, (e_tf32, e_fp32)


@pytest.mark.parametrize("act", ["relu", "tanh", "logistic"])
def test_tf32_fused_layer_passes(ann, ops, act):
    """forward (+bias+actf), data gradient (x previous derivative) and weight gradient (scale,
    beta=1 accumulate, bias column sum) of a 1024x784->2048 layer in TF32 mode vs the oracle."""
    ctx = ann.get_context()
    M, K, N = 1024, 784, 2048
    X, W, b = rnd_mat(80, M, K), rnd_mat(81, N, K, lo=-0.05, hi=0.05), rnd_mat(82, N)
    f = {"relu": A.relu, "tanh": A.antisym_logistic, "logistic": A.logistic}[act]
    d = {"relu": A.relu_der, "tanh": A.antisym_logistic_der, "logistic": A.logistic_der}[act]
    dY = rnd_mat(83, M, N)
    Xact = f(rnd_mat(84, M, K, lo=-2, hi=2))
    dW0 = rnd_mat(85, N, K)
    ctx.set_math_mode(ann.MATH_TF32)
    try:
        y = ops.linear_fwd(X, W, b, act)
        dx = ops.linear_bwd_data(dY, W, act, Xact)
        dw, db = ops.linear_bwd_weight(dY, X, scale=1.0 / 32, beta=1.0, dW0=dW0)
    finally:
        ctx.set_math_mode(ann.MATH_FP32)
    assert rel_l2(y, f((X @ W.T + b).astype(np.float32))) < TF32_TOL
    want_dx = (d(Xact) if act != "relu" else A.relu_der(Xact)) * (dY @ W)
    assert rel_l2(dx, want_dx) < TF32_TOL
    assert rel_l2(dw, dW0 + (dY.T @ X) / 32) < TF32_TOL
    assert rel_l2(db, dY.sum(axis=0) / 32) < F32_TOL * 10


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,K", [(333, 512, 200), (4096, 256, 4096), (1024, 2048, 2048), (130, 64, 33 * 4)])
def test_tf32_relu_data_gradient_mask_edges(ann, ops, M, N, K):
    """dX = (dY . W) (.) relu'(Y_below) on the tensor cores: the derivative arrives as bit masks gathered
    while the contraction runs.  Ragged rows / columns, a split-K shape, and a persistent multi-tile shape
    (earlier tiles of a CTA take the one-warp-per-quadrant path with eight masks per warp)."""
    ctx = ann.get_context()
    dY, W = rnd_mat(90, M, N), rnd_mat(91, N, K, lo=-0.05, hi=0.05)
    Y = A.relu(rnd_mat(92, M, K, lo=-1, hi=1))
    ctx.set_math_mode(ann.MATH_TF32)
    try:
        dx = ops.linear_bwd_data(dY, W, "relu", Y)
    finally:
        ctx.set_math_mode(ann.MATH_FP32)
    want = A.relu_der(Y) * (dY.astype(np.float64) @ W.astype(np.float64)).astype(np.float32)
    assert rel_l2(dx, want) < TF32_TOL
    assert np.array_equal(dx == 0, (want == 0) | (dx == 0))      # nothing leaks through a closed gate
    assert not np.any((Y <= 0) & (dx != 0))


@pytest.mark.gpu
def test_pipelined_feed_equals_train_step(ann):
    """stage() / step_staged() (two staging slots filled on a copy stream, one captured graph per slot)
    must give exactly the weights and the loss statistics of train_step() on the same bunches."""
    topo = "64 inputs 48 relu 32 tanh 10 log_softmax"

    def make():
        tr = ann.trainable.supervised_trainer(ann.mlp.all_all.generate(topo), ann.loss.multi_class_cross_entropy(), 32).build()
        tr.set_option("learning_rate", 0.05)
        tr.set_option("momentum", 0.8)
        tr.set_option("weight_decay", 1e-3)
        tr.randomize_weights(random=ann.random(77), inf=-1, sup=1, use_fanin=True, use_fanout=True)
        return tr
    a, b = make(), make()
    xs = [rnd_mat(100 + k, 32, 64) for k in range(9)]
    ts = [onehot(200 + k, 32, 10) for k in range(9)]
    la = [a.train_step(x, t)[0] for x, t in zip(xs, ts)]
    b.loss_reset()
    for x, t in zip(xs, ts):
        b.stage(x, t, 32)
        b.step_staged(32)
    mean_b, _ = b.loss_get()
    assert abs(mean_b - float(np.mean(la))) < 1e-5
    for n in a.weight_names():
        assert np.array_equal(a.weights(n), b.weights(n)), n
