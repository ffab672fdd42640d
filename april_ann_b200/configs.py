"""The five workloads of BASELINE.json (`configs`), built through the public API exactly as a user script
would (ann.mlp.all_all.generate / ann.components.*), with their shapes and algorithmic FLOP counts
(SURVEY.md 8d).  Used by bench.py, the parity tests and the profiling tools."""
import numpy as np

CONFIGS = {
    # name: topology (MLP) or "conv", loss, bunch per GPU, input size, output size
    "C1": dict(topology="256 inputs 256 tanh 128 tanh 10 log_softmax", loss="mcce", bunch=32, nin=256, nout=10,
               text="MLP 256-256-128-10 tanh/log_softmax + MCCE (the reference's digits topology, TEST/digitos/test.lua:5)"),
    "C2": dict(topology="784 inputs 2048 relu 2048 relu 10 log_softmax", loss="mcce", bunch=1024, nin=784, nout=10,
               text="MLP 784-2048-2048-10 ReLU/log_softmax + MCCE"),
    "C3": dict(topology="4096 inputs " + " ".join(["4096 tanh"] * 8), loss="mse", bunch=8192, nin=4096, nout=4096,
               text="deep MLP 8 x (4096 -> 4096 tanh) + MSE"),
    "C4": dict(topology="conv", loss="mcce", bunch=512, nin=784, nout=10,
               text="conv 5x5x16 + bias + ReLU / max-pool 2x2 / conv 5x5x32 + bias + ReLU / max-pool 2x2 / 512-256 ReLU / 10 "
                    "log_softmax + MCCE on 1x28x28 images"),
    "C5": dict(topology="512 inputs 10000 log_softmax", loss="mcce", bunch=4096, nin=512, nout=10000,
               text="NNLM-style output layer 512 -> 10000 log_softmax + MCCE"),
}
SGD_OPTIONS = (("learning_rate", 0.01), ("momentum", 0.9), ("weight_decay", 1e-4))


def workload_string(name):
    c = CONFIGS[name]
    return "%s: %s, bunch %d per GPU, SGD lr .01 momentum .9 weight decay 1e-4 (0 on biases) decay 1e-5" % (
        name, c["text"], c["bunch"])


def conv_net(ann):
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


def build_trainer(ann, name, ctx=None, bunch=None):
    """trainer of config `name` with the bench's SGD options; weights still unset."""
    c = CONFIGS[name]
    net = conv_net(ann) if c["topology"] == "conv" else ann.mlp.all_all.generate(c["topology"])
    loss = ann.loss.multi_class_cross_entropy() if c["loss"] == "mcce" else ann.loss.mse()
    tr = ann.trainable.supervised_trainer(net, loss, bunch or c["bunch"], ctx=ctx)
    tr.build(c["nin"], c["nout"]) if c["topology"] == "conv" else tr.build()
    for o, v in SGD_OPTIONS:
        tr.set_option(o, v)
    tr.set_layerwise_option("b.", "weight_decay", 0)
    return tr


def synthetic_bunch(name, seed, bunch=None):
    """inputs uniform(-1,1) (images: uniform(0,1)), targets one-hot (MSE: uniform(-1,1))  -- SURVEY.md 8d"""
    c = CONFIGS[name]
    bunch = bunch or c["bunch"]
    rng = np.random.RandomState(seed)
    lo = 0.0 if c["topology"] == "conv" else -1.0
    x = rng.uniform(lo, 1.0, size=(bunch, c["nin"])).astype(np.float32)
    if c["loss"] == "mcce":
        t = np.zeros((bunch, c["nout"]), dtype=np.float32)
        t[np.arange(bunch), rng.randint(0, c["nout"], size=bunch)] = 1.0
    else:
        t = rng.uniform(-1, 1, size=(bunch, c["nout"])).astype(np.float32)
    return x, t


def dense_layers(name):
    tok = CONFIGS[name]["topology"].split()
    sizes = [int(tok[0])] + [int(tok[i]) for i in range(2, len(tok), 2)]
    return list(zip(sizes, sizes[1:]))


def step_flops(name, bunch=None):
    """Algorithmic FLOPs per step: forward + weight gradient for every layer, data gradient for
    every layer but the first (the network-input gradient is never needed for training)."""
    bunch = bunch or CONFIGS[name]["bunch"]
    if CONFIGS[name]["topology"] == "conv":
        conv1 = 2 * bunch * 24 * 24 * 16 * 25          # 1x28x28 -> 5x5x16 -> 16x24x24
        conv2 = 2 * bunch * 8 * 8 * 32 * 400           # 16x12x12 -> 5x5x32 -> 32x8x8
        dense = 2 * bunch * (512 * 256 + 256 * 10)
        return 2 * conv1 + 3 * conv2 + 3 * dense
    layers = dense_layers(name)
    p = sum(i * o for i, o in layers)
    p1 = layers[0][0] * layers[0][1]
    return 2 * bunch * p * 2 + 2 * bunch * (p - p1)
