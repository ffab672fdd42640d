"""numpy restatement of the reference CPU training path (oracle; test infrastructure only).

All tensors are float32, row-major, batch-major ([bunch, features]); weights are
W[out, in], biases b[out, 1] -- packages/ann/ann/c_src/connection.cc:51-59.
Each class / function cites the reference lines it follows.
"""
import math
import re

import numpy as np

from .mtrand import MTRand

f32 = np.float32
NEAR_ZERO = f32(1e-6)  # packages/basics/mathcore/c_src/cmath_overloads.h:38


# ---------------------------------------------------------------------------
# scalar functors  (cmath_overloads.h)
# ---------------------------------------------------------------------------
def _sigmoid(numerator, x):
    # cmath_overloads.h:717-723  numerator / (exp(-x) + 1)
    x = np.asarray(x, dtype=f32)
    with np.errstate(over="ignore"):
        return (f32(numerator) / (np.exp(-x, dtype=f32) + f32(1.0))).astype(f32)


def logistic(x):
    return _sigmoid(1.0, x)  # :969-977


def antisym_logistic(x):
    # the reference's "tanh": 2/(1+e^-x) - 1   (:981-989)
    return (_sigmoid(2.0, x) - f32(1.0)).astype(f32)


def relu(x):
    return np.where(x > 0, x, f32(0)).astype(f32)  # :727-733


def logistic_der(y):
    v = np.clip(y, NEAR_ZERO, f32(1.0) - NEAR_ZERO).astype(f32)  # :1053-1064
    return (v * (f32(1.0) - v)).astype(f32)


def antisym_logistic_der(y):
    v = np.clip(y, f32(-1.0) + NEAR_ZERO, f32(1.0) - NEAR_ZERO).astype(f32)  # :1068-1080
    return (f32(0.5) * (f32(1.0) - v * v)).astype(f32)


def relu_der(x):
    return np.where(x > 0, f32(1), f32(0)).astype(f32)  # :1115-1124


def log_logistic(x):
    # cmath_overloads.h:993-1003: x < -10 ? x : -log1p(exp(-x))
    x = np.asarray(x, dtype=f32)
    with np.errstate(over="ignore"):
        return np.where(x < f32(-10.0), x, -np.log1p(np.exp(-x, dtype=f32), dtype=f32)).astype(f32)


def softplus(x):
    # cmath_overloads.h:1017-1023: x > 10 ? x : log1p(exp(x))
    x = np.asarray(x, dtype=f32)
    with np.errstate(over="ignore"):
        return np.where(x > f32(10.0), x, np.log1p(np.exp(x, dtype=f32), dtype=f32)).astype(f32)


def softplus_der(x):
    return logistic(x)  # m_softplus_der: from the INPUT  (:1084-1090)


def softsign(x):
    x = np.asarray(x, dtype=f32)
    return (x / (f32(1.0) + np.abs(x))).astype(f32)  # :1008-1014


def softsign_der(y):
    # :1073-1081, evaluated on the (clamped) OUTPUT
    v = np.clip(y, f32(-1.0) + NEAR_ZERO, f32(1.0) - NEAR_ZERO).astype(f32)
    aux = (f32(1.0) + np.abs(v)).astype(f32)
    return (f32(1.0) / (aux * aux)).astype(f32)


def leaky_relu(x, leak):
    return np.where(x > 0, x, f32(leak) * x).astype(f32)  # :736-744


def leaky_relu_der(x, leak):
    return np.where(x > 0, f32(1), f32(leak)).astype(f32)  # :1101-1109


def hardtanh(x, inf, sup):
    return np.clip(x, f32(inf), f32(sup)).astype(f32)  # matClamp, hardtanh_actf_component.cc:37-40


def hardtanh_der(x, inf, sup):
    # m_clamp_der :1143-1153, from the INPUT
    return np.where((x < f32(inf)) | (x > f32(sup)), f32(0), f32(1)).astype(f32)


def softmax_rows(x):
    """activation_function_kernels.cu:209-248 (CPU branch): subtract the
    (clamped) row MINIMUM, exp and sum in double, scale by float(1/sum)."""
    x = np.asarray(x, dtype=f32)
    mn = x.min(axis=1, keepdims=True)
    mx = x.max(axis=1, keepdims=True)
    mn = np.where((mx - mn) > f32(30.0), mx - f32(30.0), mn).astype(f32)
    e = np.exp((x - mn).astype(f32).astype(np.float64))
    out = e.astype(f32)
    addition = e.sum(axis=1, keepdims=True)
    ratio = (1.0 / addition).astype(f32)
    return (out * ratio).astype(f32)


def log_softmax_rows(x):
    """activation_function_kernels.cu:289-325 (CPU branch): subtract the row max
    in float, exp/sum in double, subtract float(log(sum))."""
    x = np.asarray(x, dtype=f32)
    mx = x.max(axis=1, keepdims=True)
    out = (x - mx).astype(f32)
    addition = np.exp(out.astype(np.float64)).sum(axis=1, keepdims=True)
    ratio = np.log(addition).astype(f32)
    return (out - ratio).astype(f32)


def softmax_der_rows(y, dy):
    """activation_function_kernels.cu:331-355: dX = Y * (dY - sum_j Y_j dY_j)."""
    sums = (y * dy).astype(f32).sum(axis=1, keepdims=True, dtype=f32)
    return ((dy - sums).astype(f32) * y).astype(f32)


# ---------------------------------------------------------------------------
# components  (packages/ann/ann/c_src)
# ---------------------------------------------------------------------------
class Component:
    weights_name = None

    def build(self, input_size, weights):
        """returns output size"""
        return input_size

    def forward(self, x, during_training=False):
        raise NotImplementedError

    def backprop(self, dy):
        raise NotImplementedError

    def compute_gradients(self, grads, counts):
        pass


class DotProduct(Component):
    """dot_product_component.cc:63-98 (fwd), :123-152 (bwd), :194-216 (grads)."""

    def __init__(self, input, output, weights_name):
        self.input, self.output, self.weights_name = input, output, weights_name

    def build(self, input_size, weights):
        if self.weights_name not in weights:
            weights[self.weights_name] = np.zeros((self.output, self.input), dtype=f32)
        self.w = weights
        return self.output

    def forward(self, x, during_training=False):
        self.x = x
        return (x @ self.w[self.weights_name].T).astype(f32)

    def backprop(self, dy):
        self.dy = dy
        return (dy @ self.w[self.weights_name]).astype(f32)

    def compute_gradients(self, grads, counts):
        n = self.weights_name
        grads[n] = (grads.get(n, 0) + self.dy.T @ self.x).astype(f32)
        counts[n] = counts.get(n, 0) + 1  # addToSharedCount(), :177


class Bias(Component):
    """bias_component.cc:46-73 (fwd), :87-122 (grad = column sum of dY)."""

    def __init__(self, size, weights_name):
        self.size, self.weights_name = size, weights_name

    def build(self, input_size, weights):
        if self.weights_name not in weights:
            weights[self.weights_name] = np.zeros((self.size, 1), dtype=f32)
        self.w = weights
        return self.size

    def forward(self, x, during_training=False):
        return (x + self.w[self.weights_name][:, 0][None, :]).astype(f32)

    def backprop(self, dy):
        self.dy = dy
        return dy

    def compute_gradients(self, grads, counts):
        n = self.weights_name
        g = self.dy.sum(axis=0, dtype=f32).reshape(-1, 1)
        grads[n] = (grads.get(n, 0) + g).astype(f32)
        counts[n] = counts.get(n, 0) + 1


class Actf(Component):
    """activation_function_component.cc:48-120 + the concrete *_actf_component.cc."""

    def __init__(self, kind, **params):
        self.kind = kind
        self.params = params  # leaky_relu: leak ; hardtanh: inf, sup

    def forward(self, x, during_training=False):
        self.x = x
        k = self.kind
        if k == "log_logistic":
            y = log_logistic(x)  # log_logistic_actf_component.cc:39-42
        elif k == "softplus":
            y = softplus(x)
        elif k == "softsign":
            y = softsign(x)
        elif k == "leaky_relu":
            y = leaky_relu(x, self.params.get("leak", 0.01))
        elif k == "hardtanh":
            y = hardtanh(x, self.params.get("inf", -1.0), self.params.get("sup", 1.0))
        elif k == "logistic":
            y = logistic(x)
        elif k == "tanh":
            y = antisym_logistic(x)
        elif k == "relu":
            y = relu(x)
        elif k == "linear":
            y = x.copy()
        elif k in ("softmax", "log_softmax"):
            x2 = x.reshape(x.shape[0], -1)  # flattened (softmax_actf_component.cc:38)
            y = (softmax_rows(x2) if k == "softmax" else log_softmax_rows(x2)).reshape(x.shape)
        else:
            raise ValueError(k)
        self.y = y
        return y

    def backprop(self, dy):
        k = self.kind
        if k == "logistic":
            return (logistic_der(self.y) * dy).astype(f32)  # logistic_actf_component.cc:47-54
        if k == "tanh":
            return (antisym_logistic_der(self.y) * dy).astype(f32)  # tanh_actf_component.cc:45-52
        if k == "relu":
            return (relu_der(self.x) * dy).astype(f32)  # relu_actf_component.cc:45-52
        if k == "linear":
            return dy
        if k == "softmax":
            return softmax_der_rows(self.y, dy)
        if k in ("log_softmax", "log_logistic"):
            # log_softmax_actf_component.cc:44-52, log_logistic_actf_component.cc:44-53: the derivative is
            # cancelled by the (multi-class) cross-entropy derivative -> identity
            return dy.copy()
        if k == "softplus":
            return (softplus_der(self.x) * dy).astype(f32)  # softplus_actf_component.cc:44-51
        if k == "softsign":
            return (softsign_der(self.y) * dy).astype(f32)
        if k == "leaky_relu":
            return (leaky_relu_der(self.x, self.params.get("leak", 0.01)) * dy).astype(f32)
        if k == "hardtanh":
            return (hardtanh_der(self.x, self.params.get("inf", -1.0), self.params.get("sup", 1.0)) * dy).astype(f32)
        raise ValueError(k)


class PReLU(Component):
    """prelu_actf_component.cc:55-114: y = x>0 ? x : a*x with a learnable a[size,1] (or one scalar);
    gradient da = sum over the bunch of (x<0)*x*dy; shared count += 1."""

    def __init__(self, size, weights_name, scalar=False):
        self.size, self.weights_name, self.scalar = size, weights_name, scalar

    def build(self, input_size, weights):
        self.size = self.size or input_size
        if self.weights_name not in weights:
            weights[self.weights_name] = np.zeros((1 if self.scalar else self.size, 1), dtype=f32)
        self.w = weights
        return self.size

    def _a(self):
        a = self.w[self.weights_name][:, 0]
        return a[0] if self.scalar else a[None, :]

    def forward(self, x, during_training=False):
        self.x = x
        return np.where(x > 0, x, (self._a() * x).astype(f32)).astype(f32)

    def backprop(self, dy):
        self.dy = dy
        a = self._a()
        return (np.where(self.x > 0, f32(1), a + f32(0) * self.x).astype(f32) * dy).astype(f32)

    def compute_gradients(self, grads, counts):
        n = self.weights_name
        e = (np.where(self.x < 0, f32(1), f32(0)) * self.x).astype(f32) * self.dy
        g = e.sum(dtype=f32).reshape(1, 1) if self.scalar else e.sum(axis=0, dtype=f32).reshape(-1, 1)
        grads[n] = (grads.get(n, 0) + g).astype(f32)
        counts[n] = counts.get(n, 0) + 1


class Dropout(Component):
    """dropout_component.cc:67-134: during training every unit is replaced by `value` with probability
    `prob`, the mask drawn element by element (row-major) from the component's MTRand as
    `rand() < prob`; backprop zeroes the same units; outside training the output is scaled by 1-prob
    (norm=true, the binding's default, bind_ann_base.lua.cc:1652-1668)."""

    def __init__(self, random, prob=0.5, value=0.0, norm=True):
        self.random, self.prob, self.value, self.norm = random, prob, value, norm
        self.mask = None

    def forward(self, x, during_training=False):
        if self.prob > 0.0 and (during_training or self.norm):
            if during_training:
                r = self.random.rand_array(x.size, 1.0)
                self.mask = np.where(r < float(f32(self.prob)), f32(0), f32(1)).astype(f32).reshape(x.shape)
                return np.where(self.mask < f32(0.5), f32(self.value), x).astype(f32)
            return (x * f32(1.0 - f32(self.prob))).astype(f32)
        return x

    def backprop(self, dy):
        if self.mask is not None and self.prob > 0.0:
            return np.where(self.mask < f32(0.5), f32(0), dy).astype(f32)
        return dy


class Rewrap(Component):
    """rewrap_component.cc: [bunch, prod(size)] -> [bunch, *size]."""

    def __init__(self, size):
        self.size = tuple(size)

    def build(self, input_size, weights):
        return int(np.prod(self.size))

    def forward(self, x, during_training=False):
        self.in_shape = x.shape
        return x.reshape((x.shape[0],) + self.size)

    def backprop(self, dy):
        return dy.reshape(self.in_shape)


class Flatten(Component):
    """flatten_component.cc: [bunch, ...] -> [bunch, prod]."""

    def forward(self, x, during_training=False):
        self.in_shape = x.shape
        return x.reshape(x.shape[0], -1)

    def backprop(self, dy):
        return dy.reshape(self.in_shape)


def _im2col(x, kh, kw, sh, sw):
    """[bs, C, H, W] -> [bs, oH, oW, C*kh*kw], window flattened in (plane,row,col)
    order as convolution_component.cc:108-118,175 does."""
    bs, C, H, W = x.shape
    oH, oW = (H - kh) // sh + 1, (W - kw) // sw + 1
    s0, s1, s2, s3 = x.strides
    win = np.lib.stride_tricks.as_strided(
        x, shape=(bs, oH, oW, C, kh, kw),
        strides=(s0, s2 * sh, s3 * sw, s1, s2, s3), writeable=False)
    return np.ascontiguousarray(win).reshape(bs, oH, oW, C * kh * kw), oH, oW


class Convolution(Component):
    """convolution_component.cc:135-218 (fwd: one GEMM per output pixel),
    :221-297 (bwd, beta=1 accumulation into overlapping windows),
    :300-354 (grads: one TN GEMM per pixel; shared count += #windows :302).
    kernel = (planes, kh, kw), step = (1, sh, sw); W[n, planes*kh*kw]."""

    def __init__(self, kernel, n, weights_name, step=None):
        self.kernel = tuple(kernel)
        self.step = tuple(step) if step is not None else (1,) * len(kernel)
        self.n, self.weights_name = n, weights_name

    def build(self, input_size, weights):
        ks = int(np.prod(self.kernel))
        if self.weights_name not in weights:
            weights[self.weights_name] = np.zeros((self.n, ks), dtype=f32)
        self.w = weights
        return 0  # depends on the input image; resolved at forward

    def forward(self, x, during_training=False):
        C, kh, kw = self.kernel
        _, sh, sw = self.step
        assert x.shape[1] == C
        self.x_shape = x.shape
        cols, oH, oW = _im2col(x, kh, kw, sh, sw)
        self.cols = cols
        y = (cols.reshape(-1, cols.shape[-1]) @ self.w[self.weights_name].T).astype(f32)
        bs = x.shape[0]
        return np.ascontiguousarray(y.reshape(bs, oH, oW, self.n).transpose(0, 3, 1, 2))

    def backprop(self, dy):
        C, kh, kw = self.kernel
        _, sh, sw = self.step
        bs, n, oH, oW = dy.shape
        self.dy = dy
        dyf = dy.transpose(0, 2, 3, 1).reshape(-1, n)
        dcols = (dyf @ self.w[self.weights_name]).astype(f32).reshape(bs, oH, oW, C, kh, kw)
        dx = np.zeros(self.x_shape, dtype=f32)
        for oy in range(oH):  # same pixel order as the reference's sliding window
            for ox in range(oW):
                dx[:, :, oy * sh:oy * sh + kh, ox * sw:ox * sw + kw] += dcols[:, oy, ox]
        return dx

    def compute_gradients(self, grads, counts):
        nm = self.weights_name
        bs, n, oH, oW = self.dy.shape
        dyf = self.dy.transpose(0, 2, 3, 1).reshape(-1, n)
        g = (dyf.T @ self.cols.reshape(-1, self.cols.shape[-1])).astype(f32)
        grads[nm] = (grads.get(nm, 0) + g).astype(f32)
        counts[nm] = counts.get(nm, 0) + oH * oW


class ConvolutionBias(Component):
    """convolution_bias_component.cc:120-176 (add b[plane] at every pixel),
    :181-222 (grad: sum over bunch and pixels; shared count += #windows :184)."""

    def __init__(self, n, weights_name):
        self.n, self.weights_name = n, weights_name

    def build(self, input_size, weights):
        if self.weights_name not in weights:
            weights[self.weights_name] = np.zeros((self.n, 1), dtype=f32)
        self.w = weights
        return input_size

    def forward(self, x, during_training=False):
        b = self.w[self.weights_name][:, 0]
        return (x + b[None, :, None, None]).astype(f32)

    def backprop(self, dy):
        self.dy = dy
        return dy

    def compute_gradients(self, grads, counts):
        nm = self.weights_name
        g = self.dy.sum(axis=(0, 2, 3), dtype=f32).reshape(-1, 1)
        grads[nm] = (grads.get(nm, 0) + g).astype(f32)
        counts[nm] = counts.get(nm, 0) + self.dy.shape[2] * self.dy.shape[3]


class MaxPooling(Component):
    """maxpooling_component.cc:129-206 (fwd; first maximum wins,
    matrix_ext_reductions.cu:307-320 uses a strict '>'), :208-260 (bwd scatter-add).
    kernel = (1, kh, kw); step defaults to kernel (bind_ann_base.lua.cc:1485-1487)."""

    def __init__(self, kernel, step=None):
        self.kernel = tuple(kernel)
        self.step = tuple(step) if step is not None else tuple(kernel)

    def forward(self, x, during_training=False):
        kp, kh, kw = self.kernel
        sp, sh, sw = self.step
        assert kp == 1 and sp == 1, "oracle covers per-plane pooling"
        bs, C, H, W = x.shape
        oH, oW = (H - kh) // sh + 1, (W - kw) // sw + 1
        s0, s1, s2, s3 = x.strides
        win = np.lib.stride_tricks.as_strided(
            x, shape=(bs, C, oH, oW, kh, kw),
            strides=(s0, s1, s2 * sh, s3 * sw, s2, s3), writeable=False)
        flat = win.reshape(bs, C, oH, oW, kh * kw)
        arg = flat.argmax(axis=-1)  # first maximum
        self.arg, self.x_shape = arg, x.shape
        return np.take_along_axis(flat, arg[..., None], axis=-1)[..., 0].astype(f32)

    def backprop(self, dy):
        kp, kh, kw = self.kernel
        sp, sh, sw = self.step
        bs, C, oH, oW = dy.shape
        dx = np.zeros(self.x_shape, dtype=f32)
        ay, ax = self.arg // kw, self.arg % kw
        b, c, oy, ox = np.meshgrid(np.arange(bs), np.arange(C), np.arange(oH), np.arange(oW), indexing="ij")
        np.add.at(dx, (b, c, oy * sh + ay, ox * sw + ax), dy)
        return dx


class Stack(Component):
    """stack_component.cc:81-104."""

    def __init__(self):
        self.components = []

    def push(self, c):
        self.components.append(c)
        return self

    def build(self, input_size, weights):
        size = input_size
        for c in self.components:
            size = c.build(size, weights)
        return size

    def forward(self, x, during_training=False):
        for c in self.components:
            x = c.forward(x, during_training)
        return x

    def backprop(self, dy):
        for c in reversed(self.components):
            dy = c.backprop(dy)
        return dy

    def compute_gradients(self, grads, counts):
        for c in self.components:
            c.compute_gradients(grads, counts)


def hyperplane(stack, input, output, wname, bname):
    """hyperplane_component.cc:77-97: dot_product then bias."""
    stack.push(DotProduct(input, output, wname))
    stack.push(Bias(output, bname))


def mlp_all_all(topology):
    """ann.mlp.all_all.generate -- packages/ann/ann/lua_src/annbase.lua:540-660:
    '<in> inputs <n> <actf> ...' -> stack of hyperplane(w<i>,b<i>) + actf."""
    tok = topology.split()
    assert tok[1] == "inputs"
    prev = int(tok[0])
    net = Stack()
    net.input_size = prev
    count = 1
    i = 2
    while i < len(tok):
        size, kind = int(tok[i]), tok[i + 1]
        hyperplane(net, prev, size, "w%d" % count, "b%d" % count)
        net.push(Actf(kind))
        prev = size
        count += 1
        i += 2
    return net


# ---------------------------------------------------------------------------
# losses  (packages/ann/loss/c_src)
# ---------------------------------------------------------------------------
class RunningStat:
    """packages/basics/util/c_src/mean_deviation.h:50-94 (Welford, double)."""

    def __init__(self):
        self.n = 0

    def clear(self):
        self.n = 0

    def push(self, x):
        x = float(x)
        self.n += 1
        if self.n == 1:
            self.m = x
            self.s = 0.0
        else:
            new_m = self.m + (x - self.m) / self.n
            self.s = self.s + (x - self.m) * (x - new_m)
            self.m = new_m

    def mean(self):
        return self.m if self.n > 0 else 0.0

    def variance(self):
        return self.s / (self.n - 1) if self.n > 1 else 0.0


class Loss:
    """loss_function.h:34-120 + bind_loss_functions.lua.cc:60-78."""

    def __init__(self):
        self.acc = RunningStat()

    def reset(self):
        self.acc.clear()

    def compute_loss(self, out, tgt):
        vec = self.loss_rows(out, tgt)
        mean = f32(vec.sum(dtype=f32) / f32(vec.shape[0]))  # matSum(loss)/dim, :71
        return float(mean), vec

    def accum_loss(self, vec):
        for v in vec:
            self.acc.push(v)

    def get_accum_loss(self):
        return float(f32(self.acc.mean())), float(f32(self.acc.variance()))


class MSE(Loss):
    """loss_kernels.cu:38-44,191-198 ; mse_loss_function.cc:87."""

    def loss_rows(self, o, t):
        d = (o - t).astype(f32)
        return (f32(0.5) * d * d).astype(f32).sum(axis=1, dtype=f32)

    def gradient(self, o, t):
        return (o - t).astype(f32)


class MultiClassCrossEntropy(Loss):
    """loss_kernels.cu:171-185,254-264 ; multiclass_cross_entropy_loss_function.cc:48-71.
    Input is log_softmax output."""

    def loss_rows(self, o, t):
        tc = np.clip(t, NEAR_ZERO, f32(1.0) - NEAR_ZERO).astype(f32)
        term = np.where(tc > NEAR_ZERO, -tc * o, f32(0)).astype(f32)
        return term.sum(axis=1, dtype=f32)

    def gradient(self, o, t):
        lo, hi = f32(math.log(f32(1e-6))), f32(math.log(f32(1.0) - f32(1e-6)))
        lo, hi = np.log(NEAR_ZERO).astype(f32), np.log(f32(1.0) - NEAR_ZERO).astype(f32)
        e = np.exp(np.clip(o, lo, hi).astype(f32), dtype=f32)
        return (e - t).astype(f32)


class CrossEntropy(Loss):
    """loss_kernels.cu:48-83 (loss), :121-133 (gradient); input is log_logistic output."""

    def loss_rows(self, o, t):
        le, l1e = np.log(NEAR_ZERO).astype(f32), np.log(f32(1.0) - NEAR_ZERO).astype(f32)
        log_o = np.clip(o, le, l1e).astype(f32)
        # `double o = m_exp(log_o)`: the exponential is the float one, only the log(1-o) is double
        od = np.exp(log_o, dtype=f32).astype(np.float64)
        log_inv_o = np.log(1.0 - od).astype(f32)
        tc = np.clip(t, NEAR_ZERO, f32(1.0) - NEAR_ZERO).astype(f32)
        inv_t = np.clip((f32(1.0) - t).astype(f32), NEAR_ZERO, f32(1.0) - NEAR_ZERO).astype(f32)
        s = np.where(tc > NEAR_ZERO, -tc * log_o, f32(0)).astype(f32)
        s = np.where(inv_t > NEAR_ZERO, s - inv_t * log_inv_o, s).astype(f32)
        return s.sum(axis=1, dtype=f32)

    def gradient(self, o, t):
        le, l1e = np.log(NEAR_ZERO).astype(f32), np.log(f32(1.0) - NEAR_ZERO).astype(f32)
        return (np.exp(np.clip(o, le, l1e).astype(f32), dtype=f32) - t).astype(f32)


class ZeroOne(Loss):
    """zero_one_loss_function.cc:39-132: 0/1 error per pattern; two-class (one output, threshold TH) or
    multi-class (arg-max of output vs arg-max of a dense target, or vs a [bunch,1] vector of 1-based
    class labels).  First maximum wins (matMax uses a strict '>').  Not differentiable."""

    def __init__(self, TH=0.5):
        super().__init__()
        self.TH = TH

    def loss_rows(self, o, t):
        if o.shape[1] == 1:
            pred = o[:, 0] > f32(self.TH)
            want = t[:, 0] > f32(0.5)
            return (pred != want).astype(f32)
        am = o.argmax(axis=1)
        if t.shape[1] == o.shape[1]:
            return (am != t.argmax(axis=1)).astype(f32)
        assert t.shape[1] == 1, "Incorrect target matrix bunch_size"
        return (am != (t[:, 0] - f32(1)).astype(np.int64)).astype(f32)

    def gradient(self, o, t):
        raise RuntimeError("NON DIFERENTIABLE LOSS FUNCTION")


# ---------------------------------------------------------------------------
# SGD  (packages/ann/optimizer/lua_src/optimizer_sgd.lua:50-100)
# ---------------------------------------------------------------------------
class SGD:
    DEFAULTS = dict(learning_rate=0.01, momentum=0.0, decay=1e-05, weight_decay=0.0,
                    L1_norm=0.0, max_norm_penalty=0.0)  # optimizer_sgd.lua:39-47

    def __init__(self):
        self.global_options = dict(self.DEFAULTS)
        self.layerwise = {}
        self.count = 0
        self.update = {}

    def before_eval(self, weights):
        pass

    def set_option(self, name, value):
        assert name in self.DEFAULTS
        self.global_options[name] = value

    def set_layerwise_option(self, layer, name, value):
        assert name in self.DEFAULTS
        self.layerwise.setdefault(layer, {})[name] = value

    def get_option_of(self, layer, name):
        v = self.layerwise.get(layer, {}).get(name)
        return self.global_options[name] if v is None else v  # base_optimizer.lua:111-115

    def execute(self, weights, grads):
        d0 = self.global_options["decay"]
        decay = 1.0 / (1.0 + d0 * self.count)  # :61-62
        for name, w in weights.items():
            u = self.update.get(name)
            if u is None:
                u = np.zeros_like(w)
            g = grads[name]
            lr = self.get_option_of(name, "learning_rate")
            lrd = lr * decay
            mt = self.get_option_of(name, "momentum")
            l1 = self.get_option_of(name, "L1_norm")
            l2 = self.get_option_of(name, "weight_decay")
            mnp = self.get_option_of(name, "max_norm_penalty")
            if l2 > 0.0:
                g += f32(l2) * w  # grad:axpy(l2, w)  :72
            if mt > 0.0:
                u *= f32(mt)  # :74
            else:
                u[...] = 0
            u += f32(lrd) * g  # :76
            w -= u  # :78
            if l1 > 0.0:
                _l1_truncate_gradient(w, f32(lrd * l1), u)  # :80
            if mnp > 0.0:
                _max_norm_penalty(w, mnp)  # :83
            if self.count % 100 == 0:
                _prune_subnormal(w)  # :85-87
            self.update[name] = u
        self.count += 1


def _l1_truncate_gradient(w, l1, update):
    """base_optimizer.lua:28-42."""
    z = (np.abs(w) > l1).astype(f32)
    u = (np.sign(w) * l1).astype(f32)
    w -= u
    update -= u
    w *= z


def _max_norm_penalty(w, mnp):
    """base_optimizer.lua:44-49: rows with ||row||_2 > mnp are rescaled to mnp."""
    for row in w:
        n2 = float(np.sqrt(np.dot(row.astype(f32), row.astype(f32))))
        if n2 > mnp:
            row *= f32(mnp / n2)


def _prune_subnormal(w):
    tiny = np.finfo(f32).tiny
    w[np.abs(w) < tiny] = 0
    assert np.isfinite(w).all(), "No finite number at weights matrix!!!"


class _Optimizer(SGD):
    """Shared option handling (base_optimizer.lua:88-114)."""
    DEFAULTS = {}

    def before_eval(self, weights):
        pass


class Adagrad(_Optimizer):
    """optimizer_adagrad.lua:20-83 (the reference's "adagrad" keeps an exponentially decayed mean of
    squared gradients, seeded with grad^2 at count 0)."""
    DEFAULTS = dict(learning_rate=1.0, decay=0.95, epsilon=1e-06, weight_decay=0.0, max_norm_penalty=0.0)

    def __init__(self):
        super().__init__()
        self.Egradients = {}

    def execute(self, weights, grads):
        for name, w in weights.items():
            E = self.Egradients.get(name)
            if E is None:
                E = np.zeros_like(w)
            g = grads[name]
            lr, decay, eps = (self.get_option_of(name, k) for k in ("learning_rate", "decay", "epsilon"))
            l2, mnp = self.get_option_of(name, "weight_decay"), self.get_option_of(name, "max_norm_penalty")
            if l2 > 0.0:
                g += f32(l2) * w
            if self.count == 0:
                E[...] = g * g
            else:
                E[...] = f32(decay) * E + f32(1 - decay) * (g * g)
            upd = (g * (f32(1.0) / (f32(eps) + np.sqrt(E)))).astype(f32)
            w += f32(-lr) * upd
            if mnp > 0.0:
                _max_norm_penalty(w, mnp)
            if self.count % 100 == 0:
                _prune_subnormal(w)
            self.Egradients[name] = E
        self.count += 1


class RMSProp(_Optimizer):
    """optimizer_rmsprop.lua:20-100: Nesterov look-ahead w -= mt*Eupdate BEFORE the gradient is
    evaluated; Erms = decay*Erms + (1-decay)*g^2 ; step = lr/sqrt(Erms+eps) * g (matrix:div(lr) is lr/m,
    cmath_overloads.h:1513-1519)."""
    DEFAULTS = dict(learning_rate=0.01, momentum=0.0, decay=0.99, epsilon=1e-06, weight_decay=0.0,
                    max_norm_penalty=0.0)

    def __init__(self):
        super().__init__()
        self.Eupdates, self.Erms = {}, {}

    def before_eval(self, weights):
        for name, w in weights.items():
            mt = self.get_option_of(name, "momentum")
            if mt > 0.0:
                Eu = self.Eupdates.setdefault(name, np.zeros_like(w))
                w += f32(-mt) * Eu

    def execute(self, weights, grads):
        for name, w in weights.items():
            Eu = self.Eupdates.get(name)
            if Eu is None:
                Eu = np.zeros_like(w)
            Er = self.Erms.get(name)
            if Er is None:
                Er = np.zeros_like(w)
            g = grads[name]
            lr, mt, decay, eps = (self.get_option_of(name, k) for k in ("learning_rate", "momentum", "decay", "epsilon"))
            l2, mnp = self.get_option_of(name, "weight_decay"), self.get_option_of(name, "max_norm_penalty")
            if l2 > 0.0:
                g += f32(l2) * w
            Er *= f32(decay)
            Er += f32(1 - decay) * (g * g)
            tmp = ((f32(lr) / np.sqrt(Er + f32(eps))) * g).astype(f32)
            if mt > 0.0:
                Eu *= f32(mt)
                Eu += tmp
            else:
                Eu[...] = tmp
            w -= Eu
            if mnp > 0.0:
                _max_norm_penalty(w, mnp)
            if self.count % 100 == 0:
                _prune_subnormal(w)
            self.Erms[name] = Er
            if mt > 0.0:
                self.Eupdates[name] = Eu
        self.count += 1


class Adadelta(_Optimizer):
    """optimizer_adadelta.lua:20-96."""
    DEFAULTS = dict(learning_rate=1.0, momentum=0.0, decay=0.95, epsilon=1e-06, weight_decay=0.0,
                    max_norm_penalty=0.0)

    def __init__(self):
        super().__init__()
        self.Eupdates, self.Egradients, self.update = {}, {}, {}

    def execute(self, weights, grads):
        for name, w in weights.items():
            mt = self.get_option_of(name, "momentum")
            if mt > 0.0:
                u = self.update.setdefault(name, np.zeros_like(w))
                w += f32(mt) * u
        for name, w in weights.items():
            Eu = self.Eupdates.get(name)
            if Eu is None:
                Eu = np.zeros_like(w)
            Eg = self.Egradients.get(name)
            if Eg is None:
                Eg = np.zeros_like(w)
            u = self.update.get(name)
            if u is None:
                u = np.zeros_like(w)
            g = grads[name]
            lr, decay, eps = (self.get_option_of(name, k) for k in ("learning_rate", "decay", "epsilon"))
            l2, mnp = self.get_option_of(name, "weight_decay"), self.get_option_of(name, "max_norm_penalty")
            if l2 > 0.0:
                g += f32(l2) * w
            Eg[...] = f32(decay) * Eg + f32(1 - decay) * (g * g)
            u[...] = -(g * (np.sqrt(Eu + f32(eps)) / np.sqrt(Eg + f32(eps)))).astype(f32)
            Eu[...] = f32(decay) * Eu + f32(1 - decay) * (u * u)
            w += f32(lr) * u
            if mnp > 0.0:
                _max_norm_penalty(w, mnp)
            if self.count % 100 == 0:
                _prune_subnormal(w)
            self.Eupdates[name], self.Egradients[name] = Eu, Eg
            u *= f32(lr)
            self.update[name] = u
        self.count += 1


# ---------------------------------------------------------------------------
# trainer  (packages/trainable/lua_src/supervised.lua)
# ---------------------------------------------------------------------------
class SupervisedTrainer:
    def __init__(self, net, loss, bunch_size, optimizer=None):
        self.net, self.loss, self.bunch_size = net, loss, bunch_size
        self.optimizer = optimizer or SGD()
        self.weights = {}
        self.smooth_gradients = True

    # supervised.lua:639-696
    def build(self, input=None):
        self.output_size = self.net.build(input or getattr(self.net, "input_size", 0), self.weights)
        self.weights_order = sorted(self.weights.keys())
        return self

    def set_option(self, name, value):
        self.optimizer.set_option(name, value)

    def get_option(self, name):
        return self.optimizer.global_options[name]

    # supervised.lua:262-269: Lua pattern expanded over the weight names
    def set_layerwise_option(self, pattern, name, value):
        rx = re.compile(_lua_pattern(pattern))
        for wname in self.weights_order:
            if rx.search(wname):
                self.optimizer.set_layerwise_option(wname, name, value)

    # supervised.lua:575-635 + connection.cc:77-91 (rnd_weight macro :37-46)
    def randomize_weights(self, random, inf, sup, use_fanin=False, use_fanout=False, name_match=None):
        rx = re.compile(_lua_pattern(name_match)) if name_match else None
        for wname in self.weights_order:
            if rx is not None and not rx.search(wname):
                continue
            w = self.weights[wname]
            constant = 0
            if use_fanin:
                constant += w.shape[1]
            if use_fanout:
                constant += w.shape[0]
            cinf, csup = inf, sup
            if constant > 0:
                cinf, csup = inf / math.sqrt(constant), sup / math.sqrt(constant)
            dinf, dsup = float(f32(cinf)), float(f32(csup))  # float args of randomizeWeights
            if abs(dinf) < 1e-7:
                dinf = 1e-7
            if abs(dsup) < 1e-7:
                dsup = -1e-7
            rng = dsup - dinf
            saved = (random.state.copy(), random._out.copy(), random._pos)
            vals = (random.rand_array(w.size, rng) + dinf).astype(f32)
            if (np.abs(vals) < 1e-7).any():
                # a near-zero weight is re-drawn (rnd_weight macro, connection.cc:37-46), which
                # shifts every later draw: redo this tensor with the scalar loop
                random.state, random._out, random._pos = saved
                vals = np.empty(w.size, dtype=f32)
                for i in range(w.size):
                    it = 0
                    while True:
                        v = f32(random.rand(rng) + dinf)
                        it += 1
                        if not (it < 1000 and abs(v) < 1e-7):
                            break
                    vals[i] = v
            w[...] = vals.reshape(w.shape)

    # supervised.lua:725-821
    def train_step(self, x, t, bunch_size=None, max_gradients_norm=None):
        # supervised.lua:757: `bunch_size or self.bunch_size or 1` -- NOT the row count of x
        bunch_size = bunch_size or self.bunch_size or 1
        self.optimizer.before_eval(self.weights)
        out = self.net.forward(x, True)
        tr_loss, loss_vec = self.loss.compute_loss(out, t)
        self.net.backprop(self.loss.gradient(out, t))
        grads, counts = {}, {}
        self.net.compute_gradients(grads, counts)
        if self.smooth_gradients:
            for name, g in grads.items():
                n = counts.get(name, 0) or 1
                g *= f32(1.0 / math.sqrt(n * bunch_size))  # :797-803
        if max_gradients_norm:
            # supervised.lua:805-811: matrix.dict.norm2 = sqrt(sum over tensors of norm2^2)
            nrm = math.sqrt(sum(float(np.dot(g.ravel().astype(np.float64), g.ravel().astype(np.float64)))
                                for g in grads.values()))
            if nrm > max_gradients_norm:
                ratio = f32(max_gradients_norm / nrm)
                for g in grads.values():
                    g *= ratio
        self.grads = grads
        self.optimizer.execute(self.weights, grads)
        self.loss.accum_loss(loss_vec)
        return tr_loss, loss_vec

    # supervised.lua:825-862
    def validate_step(self, x, t):
        out = self.net.forward(x, False)
        tr_loss, loss_vec = self.loss.compute_loss(out, t)
        self.loss.accum_loss(loss_vec)
        return tr_loss, loss_vec

    # supervised.lua:1149-1226 with the iterator of trainable.lua:95-330
    def train_dataset(self, input_dataset, output_dataset, shuffle=None, bunch_size=None):
        bs = bunch_size or self.bunch_size
        n = input_dataset.shape[0]
        self.loss.reset()
        idx = shuffle.shuffle(n) if shuffle is not None else list(range(n))
        for k in range(0, n, bs):
            b = idx[k:k + bs]
            self.train_step(input_dataset[b], output_dataset[b], len(b))
        return self.loss.get_accum_loss()

    def validate_dataset(self, input_dataset, output_dataset, bunch_size=None):
        bs = bunch_size or self.bunch_size
        n = input_dataset.shape[0]
        self.loss.reset()
        for k in range(0, n, bs):
            self.validate_step(input_dataset[k:k + bs], output_dataset[k:k + bs])
        return self.loss.get_accum_loss()

    # supervised.lua:1236-1245 / 1291-1430: forward only, bunch by bunch
    def calculate(self, x):
        return self.net.forward(x, False)

    def use_dataset(self, input_dataset, bunch_size=None):
        bs = bunch_size or self.bunch_size
        outs = [self.net.forward(input_dataset[k:k + bs], False) for k in range(0, input_dataset.shape[0], bs)]
        return np.concatenate(outs, axis=0)

    # supervised.lua:1556-1576: max over matching matrices of the max row 2-norm
    def norm2(self, pattern):
        rx = re.compile(_lua_pattern(pattern))
        m = 0.0
        for name in self.weights_order:
            if rx.search(name):
                w = self.weights[name].astype(np.float64)
                m = max(m, float(np.sqrt((w * w).sum(axis=1)).max()))
        return m


def _lua_pattern(p):
    """The handful of Lua patterns the reference scripts use ('b.', '.*w.*', 'w1')
    mean the same thing as Python regexes."""
    return p.replace("%", "\\")
