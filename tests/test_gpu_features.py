"""GPU parity of the SURVEY.md 8(f) rows and of the trainer behaviours the first round left open:
extra activation functions, PReLU, dropout (exact MT19937 mask stream), zero-one loss, use_dataset,
max_gradients_norm, adagrad / rmsprop / adadelta, checkpoint + resume of the optimizer state, and the
robustness cases of ADVICE.md (options changed after a step graph was captured, scratch growth between
captured graphs, two contexts in one process).  Everything through the C ABI, against the oracle.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import MTRand  # noqa: E402
from oracle import april as A  # noqa: E402

F32_TOL = 1e-5


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


def rnd(seed, *shape, lo=-1.0, hi=1.0):
    return np.random.RandomState(seed).uniform(lo, hi, size=shape).astype(np.float32)


def onehot(seed, M, C):
    t = np.zeros((M, C), dtype=np.float32)
    t[np.arange(M), np.random.RandomState(seed).randint(0, C, size=M)] = 1.0
    return t


def sync_weights(tr, ref):
    for n in tr.weight_names():
        tr.set_weights(n, ref.weights[n])


def assert_same_state(tr, ref, tol=2e-5):
    for n in tr.weight_names():
        assert rel_l2(tr.weights(n), ref.weights[n]) < tol, n


# ------------------------------------------------------------------ activation functions
@pytest.mark.parametrize("kind,params", [("log_logistic", {}), ("softplus", {}), ("softsign", {}),
                                         ("leaky_relu", {"leak": 0.2}), ("hardtanh", {"inf": -0.5, "sup": 0.75})])
def test_extra_activations_forward_and_derivative(ops, kind, params):
    """activation_function_kernels.cu:52-182 / cmath_overloads.h functors, incl. the asymptotic branches
    (|x| > 10) of log_logistic / softplus and the clamp edges of hardtanh."""
    x = rnd(1, 41, 67, lo=-14, hi=14)
    x[0, :4] = [-0.5, 0.75, 0.0, 10.0]
    a = A.Actf(kind, **params)
    y_ref = a.forward(x)
    y = ops.actf_fwd_ex(kind, x, **params)
    assert np.allclose(y, y_ref, rtol=2e-6, atol=2e-7)
    dy = rnd(2, 41, 67)
    dx = ops.actf_bwd_ex(kind, x, y_ref, dy, **params)
    assert np.allclose(dx, a.backprop(dy), rtol=2e-6, atol=2e-7)


@pytest.mark.parametrize("topo,loss", [("20 inputs 16 softplus 12 softsign 5 log_softmax", "mcce"),
                                       ("20 inputs 16 leaky_relu 12 hardtanh 5 log_softmax", "mcce"),
                                       ("20 inputs 16 tanh 6 log_logistic", "ce")])
def test_nets_with_extra_activations_train_like_the_oracle(ann, topo, loss):
    """log_logistic makes cross_entropy trainable (log_logistic_actf_component.cc + cross_entropy_loss_function.cc)."""
    ref_loss = A.MultiClassCrossEntropy() if loss == "mcce" else A.CrossEntropy()
    gpu_loss = ann.loss.multi_class_cross_entropy() if loss == "mcce" else ann.loss.cross_entropy()
    ref = A.SupervisedTrainer(A.mlp_all_all(topo), ref_loss, 9).build()
    tr = ann.trainable.supervised_trainer(ann.mlp.all_all.generate(topo), gpu_loss, 9).build()
    for o, v in (("learning_rate", 0.05), ("momentum", 0.5), ("weight_decay", 1e-3)):
        ref.set_option(o, v)
        tr.set_option(o, v)
    ref.randomize_weights(random=MTRand(7), inf=-1, sup=1, use_fanin=True)
    tr.randomize_weights(random=ann.random(7), inf=-1, sup=1, use_fanin=True)
    tr.set_flag("keep_gradients", 1)
    nout = int(topo.split()[-2])
    x = rnd(3, 9, 20)
    t = onehot(4, 9, nout) if loss == "mcce" else (rnd(5, 9, nout, lo=0, hi=1) > 0.5).astype(np.float32)
    for step in range(4):
        l_gpu, rows_gpu = tr.train_step(x, t)
        l_ref, rows_ref = ref.train_step(x, t)
        assert abs(l_gpu - l_ref) <= 5e-6 * max(1, abs(l_ref)), (step, l_gpu, l_ref)
        for n in tr.weight_names():
            assert rel_l2(tr.gradients(n), ref.grads[n]) < 3e-5, (step, n)
            assert rel_l2(tr.weights(n), ref.weights[n]) < F32_TOL, (step, n)


@pytest.mark.parametrize("scalar", [False, True])
def test_prelu_component(ann, scalar):
    """prelu_actf_component.cc: learnable slope (per unit, or one scalar), its gradient and shared count."""
    c = ann.components
    net = c.stack().push(c.hyperplane(input=12, output=10, name="h1", dot_product_name="w1", bias_name="b1",
                                      dot_product_weights="w1", bias_weights="b1"),
                         c.actf.prelu(size=10, scalar=scalar, name="prelu", weights="a1"),
                         c.hyperplane(input=10, output=4, name="h2", dot_product_name="w2", bias_name="b2",
                                      dot_product_weights="w2", bias_weights="b2"),
                         c.actf.log_softmax())
    onet = A.Stack()
    A.hyperplane(onet, 12, 10, "w1", "b1")
    onet.push(A.PReLU(10, "a1", scalar=scalar))
    A.hyperplane(onet, 10, 4, "w2", "b2")
    onet.push(A.Actf("log_softmax"))
    onet.input_size = 12
    ref = A.SupervisedTrainer(onet, A.MultiClassCrossEntropy(), 7).build()
    tr = ann.trainable.supervised_trainer(net, ann.loss.multi_class_cross_entropy(), 7).build()
    assert sorted(tr.weight_names()) == sorted(ref.weights)
    ref.randomize_weights(random=MTRand(11), inf=-1, sup=1, use_fanin=True)
    ref.weights["a1"][...] = 0.25
    sync_weights(tr, ref)
    tr.set_flag("keep_gradients", 1)
    for o, v in (("learning_rate", 0.1), ("momentum", 0.3)):
        ref.set_option(o, v)
        tr.set_option(o, v)
    x, t = rnd(6, 7, 12), onehot(7, 7, 4)
    for step in range(3):
        l_gpu, _ = tr.train_step(x, t)
        l_ref, _ = ref.train_step(x, t)
        assert abs(l_gpu - l_ref) < 5e-6
        for n in tr.weight_names():
            assert rel_l2(tr.gradients(n), ref.grads[n]) < 3e-5, (step, n)
            assert rel_l2(tr.weights(n), ref.weights[n]) < F32_TOL, (step, n)


# ------------------------------------------------------------------ dropout
def test_dropout_mask_follows_the_reference_stream(ops):
    """dropout_component.cc:91-95: element i is dropped iff the i-th rand() of the MT19937 is < prob.  Lengths
    around the 624-word block and the 227-word wave of the device generator; consecutive calls continue
    the same stream."""
    from april_ann_b200.ops import DropoutStream
    st = DropoutStream(seed=4321)
    want = MTRand(4321)
    for n in (1, 5, 227, 228, 623, 624, 625, 5000, 70001):
        got = st.mask(n, prob=0.3)
        r = want.rand_array(n, 1.0)
        ref = np.where(r < float(np.float32(0.3)), 0.0, 1.0).astype(np.float32)
        assert np.array_equal(got, ref), n


def test_dropout_component_trains_like_the_oracle(ann):
    c = ann.components
    net = c.stack().push(c.hyperplane(input=30, output=24, name="h1", dot_product_name="w1", bias_name="b1",
                                      dot_product_weights="w1", bias_weights="b1"),
                         c.actf.tanh(name="a1"),
                         c.dropout(random=ann.random(99), prob=0.4, name="drop"),
                         c.hyperplane(input=24, output=5, name="h2", dot_product_name="w2", bias_name="b2",
                                      dot_product_weights="w2", bias_weights="b2"),
                         c.actf.log_softmax())
    onet = A.Stack()
    A.hyperplane(onet, 30, 24, "w1", "b1")
    onet.push(A.Actf("tanh")).push(A.Dropout(MTRand(99), prob=0.4))
    A.hyperplane(onet, 24, 5, "w2", "b2")
    onet.push(A.Actf("log_softmax"))
    onet.input_size = 30
    ref = A.SupervisedTrainer(onet, A.MultiClassCrossEntropy(), 11).build()
    tr = ann.trainable.supervised_trainer(net, ann.loss.multi_class_cross_entropy(), 11).build()
    ref.randomize_weights(random=MTRand(5), inf=-1, sup=1, use_fanin=True)
    sync_weights(tr, ref)
    for o, v in (("learning_rate", 0.1), ("momentum", 0.5)):
        ref.set_option(o, v)
        tr.set_option(o, v)
    x, t = rnd(8, 11, 30), onehot(9, 11, 5)
    for step in range(5):      # step 3 onwards replays the captured graph: the mask stream must keep advancing
        l_gpu, _ = tr.train_step(x, t)
        l_ref, _ = ref.train_step(x, t)
        assert abs(l_gpu - l_ref) < 5e-6 * max(1, abs(l_ref)), (step, l_gpu, l_ref)
    assert_same_state(tr, ref)
    # outside training the output is scaled by 1 - prob (norm = true)
    y = tr.calculate(x)
    assert rel_l2(y, ref.calculate(x)) < F32_TOL


# ------------------------------------------------------------------ zero-one loss / use_dataset
def test_zero_one_loss(ann, ops):
    """ann/loss/c_src/zero_one_loss_function.cc:39-132 (same cases as packages/ann/loss/test/test.lua's zero-one
    block): two-class with threshold, multi-class vs dense target, multi-class vs 1-based labels; ties go to
    the first maximum."""
    o = rnd(10, 50, 7)
    o[0, :] = 0.3                   # tie: arg-max 0
    t = onehot(11, 50, 7)
    z = A.ZeroOne()
    assert np.array_equal(ops.zero_one_loss(o, t), z.loss_rows(o, t))
    labels = (t.argmax(axis=1) + 1).astype(np.float32).reshape(-1, 1)
    assert np.array_equal(ops.zero_one_loss(o, labels), z.loss_rows(o, labels))
    o1, t1 = rnd(12, 40, 1, lo=0, hi=1), (rnd(13, 40, 1, lo=0, hi=1) > 0.5).astype(np.float32)
    assert np.array_equal(ops.zero_one_loss(o1, t1, TH=0.6), A.ZeroOne(0.6).loss_rows(o1, t1))
    # through the trainer: validate_dataset returns the error rate; training with it is refused
    topo = "20 inputs 16 tanh 7 log_softmax"
    tr = ann.trainable.supervised_trainer(ann.mlp.all_all.generate(topo), ann.loss.zero_one(), 16).build()
    ref = A.SupervisedTrainer(A.mlp_all_all(topo), A.ZeroOne(), 16).build()
    ref.randomize_weights(random=MTRand(3), inf=-1, sup=1, use_fanin=True)
    sync_weights(tr, ref)
    x = rnd(14, 50, 20)
    m_gpu, v_gpu = tr.validate_dataset(x, t)
    m_ref, v_ref = ref.validate_dataset(x, t)
    assert abs(m_gpu - m_ref) < 1e-6 and abs(v_gpu - v_ref) < 1e-5
    with pytest.raises(ann.B200Error) as e:
        tr.train_step(x[:16], t[:16])
    assert e.value.code == 128


def test_use_dataset(ann):
    """supervised.lua:1291-1430: forward over a dataset whose length is not a multiple of the bunch."""
    topo = "20 inputs 16 relu 5 softmax"
    tr = ann.trainable.supervised_trainer(ann.mlp.all_all.generate(topo), ann.loss.mse(), 16).build()
    ref = A.SupervisedTrainer(A.mlp_all_all(topo), A.MSE(), 16).build()
    ref.randomize_weights(random=MTRand(3), inf=-1, sup=1, use_fanin=True)
    sync_weights(tr, ref)
    x = rnd(15, 53, 20)
    y = tr.use_dataset(x)
    assert y.shape == (53, 5) and rel_l2(y, ref.use_dataset(x)) < F32_TOL


# ------------------------------------------------------------------ train_step arguments
def test_train_step_bunch_size_semantics_and_gradient_clip(ann):
    """supervised.lua:757,800: the smoothing factor uses the trainer's bunch_size unless one is passed, not the
    row count of the bunch; :805-811: global gradient-norm clip."""
    topo = "20 inputs 16 tanh 5 log_softmax"
    ref = A.SupervisedTrainer(A.mlp_all_all(topo), A.MultiClassCrossEntropy(), 8).build()
    tr = ann.trainable.supervised_trainer(ann.mlp.all_all.generate(topo), ann.loss.multi_class_cross_entropy(), 8).build()
    ref.randomize_weights(random=MTRand(21), inf=-1, sup=1, use_fanin=True)
    sync_weights(tr, ref)
    tr.set_flag("keep_gradients", 1)
    for o, v in (("learning_rate", 0.2), ("momentum", 0.4), ("weight_decay", 1e-3)):
        ref.set_option(o, v)
        tr.set_option(o, v)
    x, t = rnd(16, 5, 20), onehot(17, 5, 5)      # 5 rows, trainer bunch 8
    tr.train_step(x, t)
    ref.train_step(x, t)
    for n in tr.weight_names():
        assert rel_l2(tr.gradients(n), ref.grads[n]) < 2e-5, n
    tr.train_step(x, t, bunch_size=5)
    ref.train_step(x, t, bunch_size=5)
    for n in tr.weight_names():
        assert rel_l2(tr.gradients(n), ref.grads[n]) < 2e-5, n
    # clip: a threshold below the current norm (active) and far above it (inactive)
    for thr in (0.05, 0.05, 50.0, 0.05):
        tr.train_step(x, t, max_gradients_norm=thr)
        ref.train_step(x, t, max_gradients_norm=thr)
        total = np.sqrt(sum(float((tr.gradients(n).astype(np.float64) ** 2).sum()) for n in tr.weight_names()))
        # (with weight decay the written-back gradient carries +l2*w after the clip, so only a loose bound)
        assert total < max(thr, 0.0) * 1.2 + 1.0
        for n in tr.weight_names():
            assert rel_l2(tr.gradients(n), ref.grads[n]) < 3e-5, (thr, n)
    assert_same_state(tr, ref)


# ------------------------------------------------------------------ other optimizers
@pytest.mark.parametrize("name,opts", [
    ("adagrad", {"learning_rate": 0.05, "weight_decay": 1e-3}),
    ("rmsprop", {"learning_rate": 0.01, "momentum": 0.5, "weight_decay": 1e-3}),
    ("rmsprop", {"learning_rate": 0.01}),
    ("adadelta", {"learning_rate": 1.0, "momentum": 0.3, "weight_decay": 1e-3, "max_norm_penalty": 1.5}),
])
def test_other_optimizers_match_the_oracle(ann, name, opts):
    """optimizer_{adagrad,rmsprop,adadelta}.lua incl. rmsprop's Nesterov look-ahead before the gradient and
    adadelta's momentum on the previous update."""
    topo = "20 inputs 16 tanh 12 relu 5 log_softmax"
    oopt = {"adagrad": A.Adagrad, "rmsprop": A.RMSProp, "adadelta": A.Adadelta}[name]()
    ref = A.SupervisedTrainer(A.mlp_all_all(topo), A.MultiClassCrossEntropy(), 9, optimizer=oopt).build()
    tr = ann.trainable.supervised_trainer(ann.mlp.all_all.generate(topo), ann.loss.multi_class_cross_entropy(), 9,
                                          optimizer=getattr(ann.optimizer, name)()).build()
    for k in oopt.DEFAULTS:
        assert tr.get_option(k) == oopt.DEFAULTS[k], k
    for k, v in opts.items():
        ref.set_option(k, v)
        tr.set_option(k, v)
    ref.randomize_weights(random=MTRand(31), inf=-1, sup=1, use_fanin=True)
    sync_weights(tr, ref)
    xs = [rnd(40 + i, 9, 20) for i in range(3)]
    ts = [onehot(50 + i, 9, 5) for i in range(3)]
    for step in range(6):
        l_gpu, _ = tr.train_step(xs[step % 3], ts[step % 3])
        l_ref, _ = ref.train_step(xs[step % 3], ts[step % 3])
        assert abs(l_gpu - l_ref) < 2e-5 * max(1, abs(l_ref)), (step, l_gpu, l_ref)
        for n in tr.weight_names():
            assert rel_l2(tr.weights(n), ref.weights[n]) < 5e-5, (step, n)
    with pytest.raises(ann.B200Error):
        tr.set_option("L1_norm", 0.1)       # not an option of these optimizers


# ------------------------------------------------------------------ checkpoint / resume
@pytest.mark.parametrize("name", ["sgd", "adadelta"])
def test_checkpoint_roundtrip_continues_bit_exactly(ann, name):
    """optimizer_sgd.lua:102-119 exports options + count + update: a run resumed from (weights, optimizer state,
    count) must continue exactly like the uninterrupted one (the decayed learning rate depends on count)."""
    topo = "24 inputs 32 tanh 16 relu 6 log_softmax"

    def make():
        tr = ann.trainable.supervised_trainer(ann.mlp.all_all.generate(topo), ann.loss.multi_class_cross_entropy(), 12,
                                              optimizer=getattr(ann.optimizer, name)()).build()
        if name == "sgd":
            for o, v in (("learning_rate", 0.1), ("momentum", 0.9), ("weight_decay", 1e-3), ("decay", 0.05)):
                tr.set_option(o, v)
        else:
            tr.set_option("momentum", 0.2)
        tr.randomize_weights(random=ann.random(77), inf=-1, sup=1, use_fanin=True)
        return tr
    xs = [rnd(60 + i, 12, 24) for i in range(8)]
    ts = [onehot(70 + i, 12, 6) for i in range(8)]
    a = make()
    for i in range(4):
        a.train_step(xs[i], ts[i])
    sd = a.state_dict()
    assert sd["count"] == 4
    for i in range(4, 8):
        a.train_step(xs[i], ts[i])
    b = make()
    b.load_state_dict(sd)
    assert b.get_count() == 4
    for i in range(4, 8):
        b.train_step(xs[i], ts[i])
    for n in a.weight_names():
        assert np.array_equal(a.weights(n), b.weights(n)), n
        assert np.array_equal(a.updates(n), b.updates(n)), n
    assert a.get_count() == b.get_count() == 8
    # and the count really matters: resuming with count 0 must NOT reproduce the run (sgd lr decay 0.05)
    if name == "sgd":
        c = make()
        c.load_state_dict(dict(sd, count=0))
        for i in range(4, 8):
            c.train_step(xs[i], ts[i])
        assert not np.array_equal(a.weights("w1"), c.weights("w1"))


# ------------------------------------------------------------------ robustness
def test_options_changed_after_graph_capture_take_effect(ann):
    """Captured step graphs bake kernel arguments in: decay, the gradient scale, the write-back flag, the
    topology.  Changing them after the graph exists must still change the next step."""
    topo = "20 inputs 16 tanh 5 log_softmax"
    ref = A.SupervisedTrainer(A.mlp_all_all(topo), A.MultiClassCrossEntropy(), 9).build()
    tr = ann.trainable.supervised_trainer(ann.mlp.all_all.generate(topo), ann.loss.multi_class_cross_entropy(), 9).build()
    ref.randomize_weights(random=MTRand(41), inf=-1, sup=1, use_fanin=True)
    sync_weights(tr, ref)
    x, t = rnd(80, 9, 20), onehot(81, 9, 5)
    for _ in range(4):                       # graph captured at the second step, replayed afterwards
        tr.train_step(x, t)
        ref.train_step(x, t)
    assert_same_state(tr, ref)
    for o, v in (("decay", 0.5), ("learning_rate", 0.3), ("momentum", 0.7), ("weight_decay", 1e-2)):
        tr.set_option(o, v)
        ref.set_option(o, v)
        for _ in range(2):
            tr.train_step(x, t)
            ref.train_step(x, t)
        assert_same_state(tr, ref)
    tr.set_flag("smooth_gradients", 0)
    ref.smooth_gradients = False
    for _ in range(3):
        tr.train_step(x, t)
        ref.train_step(x, t)
    assert_same_state(tr, ref)
    tr.set_flag("keep_gradients", 1)
    tr.train_step(x, t)
    ref.train_step(x, t)
    for n in tr.weight_names():
        assert rel_l2(tr.gradients(n), ref.grads[n]) < 3e-5, n


def test_scratch_growth_between_captured_graphs(ann):
    """A small bunch is captured first, a larger one (which outgrows the per-branch scratch and the pooling
    index block) afterwards, then the first graph is replayed: nothing may read or write a freed block."""
    from test_gpu_fullsize import c4_gpu, c4_oracle
    ref = A.SupervisedTrainer(c4_oracle(), A.MultiClassCrossEntropy(), 64).build(784)
    tr = ann.trainable.supervised_trainer(c4_gpu(ann), ann.loss.multi_class_cross_entropy(10), 64).build(784, 10)
    for o, v in (("learning_rate", 0.01), ("momentum", 0.5)):
        ref.set_option(o, v)
        tr.set_option(o, v)
    ref.randomize_weights(random=MTRand(1234), inf=-1, sup=1, use_fanin=True, use_fanout=True)
    sync_weights(tr, ref)
    xs, ts = rnd(90, 4, 784, lo=0, hi=1), onehot(91, 4, 10)
    xl, tl = rnd(92, 64, 784, lo=0, hi=1), onehot(93, 64, 10)
    for (x, t, b) in [(xs, ts, 4)] * 3 + [(xl, tl, 64)] * 3 + [(xs, ts, 4)] * 3 + [(xl, tl, 64)] * 2:
        l_gpu, _ = tr.train_step(x, t, bunch_size=b)
        l_ref, _ = ref.train_step(x, t, bunch_size=b)
        assert abs(l_gpu - l_ref) < 2e-5 * max(1, abs(l_ref))
    assert_same_state(tr, ref, tol=5e-5)


def test_many_small_tensors_bucket_plan(ann):
    """A deep net of small layers has more than 32 tensors in one 4 MB bucket: single-device training must work
    (and the replica-group planner is exercised on CPU in tests/test_data_parallel_cpu.py)."""
    topo = "16 inputs " + " ".join(["16 tanh"] * 20) + " 4 log_softmax"
    ref = A.SupervisedTrainer(A.mlp_all_all(topo), A.MultiClassCrossEntropy(), 8).build()
    tr = ann.trainable.supervised_trainer(ann.mlp.all_all.generate(topo), ann.loss.multi_class_cross_entropy(), 8).build()
    ref.randomize_weights(random=MTRand(61), inf=-1, sup=1, use_fanin=True)
    sync_weights(tr, ref)
    x, t = rnd(94, 8, 16), onehot(95, 8, 4)
    for _ in range(3):
        tr.train_step(x, t)
        ref.train_step(x, t)
    assert len(tr.weight_names()) == 42
    assert_same_state(tr, ref)


def test_two_contexts_in_one_process(ann):
    """one host thread may drive several contexts.  Two contexts (on this box both on device 0;
    with more devices visible the second one goes to device 1), interleaved steps of a net whose kernels need
    more than 48 KB of dynamic shared memory, both in the tensor-core mode."""
    ndev = 2 if ann.is_cuda_available() and _device_count(ann) > 1 else 1
    ctxs = [ann.get_context(), ann.Context(ndev - 1)]
    import gc
    topo = "784 inputs 512 tanh 256 tanh 10 log_softmax"
    trs, refs = [], []
    for i, ctx in enumerate(ctxs):
        ctx.set_math_mode(ann.MATH_TF32)
        tr = ann.trainable.supervised_trainer(ann.mlp.all_all.generate(topo), ann.loss.multi_class_cross_entropy(), 256,
                                              ctx=ctx).build()
        ref = A.SupervisedTrainer(A.mlp_all_all(topo), A.MultiClassCrossEntropy(), 256).build()
        ref.randomize_weights(random=MTRand(100 + i), inf=-1, sup=1, use_fanin=True, use_fanout=True)
        sync_weights(tr, ref)
        for o, v in (("learning_rate", 0.05), ("momentum", 0.9)):
            tr.set_option(o, v)
            ref.set_option(o, v)
        trs.append(tr)
        refs.append(ref)
    try:
        for step in range(4):
            for i in (0, 1):
                x, t = rnd(200 + 10 * i + step, 256, 784), onehot(300 + 10 * i + step, 256, 10)
                l_gpu, _ = trs[i].train_step(x, t)
                l_ref, _ = refs[i].train_step(x, t)
                assert abs(l_gpu - l_ref) < 2e-3 * max(1, abs(l_ref)), (step, i)
        for i in (0, 1):
            for n in trs[i].weight_names():
                assert rel_l2(trs[i].weights(n), refs[i].weights[n]) < 2e-3, (i, n)
    finally:
        ann.get_context().set_math_mode(ann.MATH_FP32)
        tr = None
        trs.clear()
        gc.collect()          # every trainer of the second context is gone before the context is destroyed
        ctxs[1].close()


def _device_count(ann):
    import ctypes as C
    from april_ann_b200._lib import lib
    n = C.c_int(0)
    lib.b200_device_count(C.byref(n))
    return n.value
