"""Oracle-side builders of the BASELINE.json workloads (test infrastructure only: bench.py's cpu_baseline /
--impl reference / parity-check legs and tests/)."""
from . import april as A

TOPOLOGIES = {
    "C1": "256 inputs 256 tanh 128 tanh 10 log_softmax",
    "C2": "784 inputs 2048 relu 2048 relu 10 log_softmax",
    "C3": "4096 inputs " + " ".join(["4096 tanh"] * 8),
    "C5": "512 inputs 10000 log_softmax",
}


def conv_net():
    """The C4 stack of SURVEY.md 8d, component by component as packages/ann/ann/test/test-convolution-digits.lua:69-105
    builds its own."""
    net = A.Stack()
    net.push(A.Rewrap((1, 28, 28)))
    net.push(A.Convolution((1, 5, 5), 16, "w1")).push(A.ConvolutionBias(16, "b1")).push(A.Actf("relu"))
    net.push(A.MaxPooling((1, 2, 2)))
    net.push(A.Convolution((16, 5, 5), 32, "w2")).push(A.ConvolutionBias(32, "b2")).push(A.Actf("relu"))
    net.push(A.MaxPooling((1, 2, 2)))
    net.push(A.Flatten())
    A.hyperplane(net, 512, 256, "w3", "b3")
    net.push(A.Actf("relu"))
    A.hyperplane(net, 256, 10, "w4", "b4")
    net.push(A.Actf("log_softmax"))
    net.input_size = 784
    return net


def build_trainer(name, bunch):
    net = conv_net() if name == "C4" else A.mlp_all_all(TOPOLOGIES[name])
    loss = A.MSE() if name == "C3" else A.MultiClassCrossEntropy()
    tr = A.SupervisedTrainer(net, loss, bunch).build(784 if name == "C4" else None)
    for o, v in (("learning_rate", 0.01), ("momentum", 0.9), ("weight_decay", 1e-4)):
        tr.set_option(o, v)
    tr.set_layerwise_option("b.", "weight_decay", 0)
    return tr
