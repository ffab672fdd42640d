"""april_ann_b200 -- B200-native (sm_100a) training hot path behind the APRIL-ANN API.

Python mirror of the reference's Lua surface for the path (same names, argument meaning
and error behaviour), over the C ABI in include/b200ann.h / b200ann_host.h:

    ann.mlp.all_all.generate        packages/ann/ann/lua_src/annbase.lua:509-660
    ann.components.*                packages/ann/ann/binding/bind_ann_base.lua.cc:287-2187
    ann.loss.*                      packages/ann/loss/binding/bind_loss_functions.lua.cc:50-140
    trainable.supervised_trainer    packages/trainable/lua_src/supervised.lua
    random                          packages/basics/random/binding/bind_mtrand.lua.cc

Everything numeric runs in hand-written CUDA kernels (libb200ann.so); there is no CPU
fallback and no dependency on the test oracle.
"""
import ctypes as C

import numpy as np

from ._lib import B200Error, lib, check, nonnull, fptr, b  # noqa: F401

MATH_FP32, MATH_TF32 = 0, 1
_f32 = np.float32


def _as_f32(a):
    return np.ascontiguousarray(a, dtype=_f32)


# ----------------------------------------------------------------------------- context
class Context:
    """Per-device runtime (streams, caching pool).  Replaces the process-global GPUHelper
    of the reference (mathcore/c_src/gpu_helper.h:42-148)."""

    def __init__(self, device=0):
        h = C.c_void_p()
        check(lib.b200_create(C.c_int(device), C.byref(h)))
        self.h = h
        self.device = device

    def close(self):
        if self.h:
            lib.b200_destroy(self.h)
            self.h = None

    def set_math_mode(self, mode):
        check(lib.b200_set_math_mode(self.h, C.c_int(mode)))

    def sync(self):
        check(lib.b200_sync(self.h))

    def launch_count(self):
        n = C.c_uint64()
        check(lib.b200_launch_count(self.h, C.byref(n)))
        return n.value

    def sm_count(self):
        n = C.c_int()
        check(lib.b200_sm_count(self.h, C.byref(n)))
        return n.value


_default_ctx = None


def get_context(device=None):
    """mathcore.set_use_cuda_default(true) analogue: the default device context."""
    global _default_ctx
    if _default_ctx is None:
        import os
        _default_ctx = Context(int(os.environ.get("LOCAL_RANK", "0")) if device is None else device)
    return _default_ctx


def is_cuda_available():
    """util.is_cuda_available()  (basics/util/binding/bind_util.lua.cc:160)"""
    n = C.c_int(0)
    return lib.b200_device_count(C.byref(n)) == 0 and n.value > 0


# ----------------------------------------------------------------------------- random
class random:  # noqa: N801  (the reference's Lua class is lower-case)
    """random(seed): MT19937 with the reference's accessors."""

    def __init__(self, seed):
        self.h = nonnull(lib.b200h_random_new(C.c_uint32(seed)))

    def __del__(self):
        # (at interpreter shutdown the module globals may already be gone)
        if getattr(self, "h", None) and lib is not None:
            lib.b200h_random_free(self.h)
            self.h = None

    def rand(self, n=1.0):
        return lib.b200h_random_rand(self.h, C.c_double(n))

    def randInt(self, x=None, y=None):  # noqa: N802
        if x is None:
            return lib.b200h_random_randint(self.h, C.c_uint32(0xFFFFFFFF))
        if y is None:
            return lib.b200h_random_randint(self.h, C.c_uint32(x))
        if y < x:
            raise B200Error(128, "first argument must be <= second argument")
        return x + lib.b200h_random_randint(self.h, C.c_uint32(y - x))

    def shuffle(self, size):
        """0-based permutation (the Lua binding returns the same permutation 1-based)."""
        out = np.empty(size, dtype=np.int32)
        check(lib.b200h_random_shuffle(self.h, C.c_int(size), out.ctypes.data_as(C.POINTER(C.c_int))))
        return out


# ----------------------------------------------------------------------------- components
class _Component:
    def __init__(self, handle):
        self.h = nonnull(handle)

    def __del__(self):
        # (at interpreter shutdown the module globals may already be gone)
        if getattr(self, "h", None) and lib is not None:
            lib.b200h_component_free(self.h)
            self.h = None


class _Stack(_Component):
    def __init__(self, name="stack", input=None):  # noqa: A002
        super().__init__(lib.b200h_stack_new(b(name)))
        self._children = []

    def push(self, *comps):
        for c in comps:
            check(lib.b200h_stack_push(self.h, c.h))
            self._children.append(c)
        return self


def _ints(v):
    arr = (C.c_int * len(v))(*[int(x) for x in v])
    return arr


class _Actf:
    @staticmethod
    def _make(kind):
        def ctor(name=None, **_):
            return _Component(lib.b200h_actf_new(b(kind), b(name or "")))
        return ctor


for _k in ("logistic", "tanh", "relu", "softmax", "log_softmax", "linear", "log_logistic", "softplus", "softsign"):
    setattr(_Actf, _k, staticmethod(_Actf._make(_k)))


def _leaky_relu(name=None, leak=0.01, **_):
    """ann.components.actf.leaky_relu{ leak= }  (leaky_relu_actf_component.cc)"""
    return _Component(lib.b200h_actf_new_ex(b("leaky_relu"), b(name or ""), C.c_float(leak), C.c_float(0.0)))


def _hardtanh(name=None, inf=-1.0, sup=1.0, **_):
    """ann.components.actf.hardtanh{ inf=, sup= }  (hardtanh_actf_component.cc)"""
    return _Component(lib.b200h_actf_new_ex(b("hardtanh"), b(name or ""), C.c_float(inf), C.c_float(sup)))


def _prelu(size=0, scalar=False, name="prelu", weights=None, **_):
    """ann.components.actf.prelu{ size=, scalar=, name=, weights= }  (prelu_actf_component.cc)"""
    return _Component(lib.b200h_prelu_new(b(name), b(weights or name), C.c_uint(size), C.c_int(bool(scalar))))


_Actf.leaky_relu = staticmethod(_leaky_relu)
_Actf.hardtanh = staticmethod(_hardtanh)
_Actf.prelu = staticmethod(_prelu)


class components:  # noqa: N801
    actf = _Actf

    @staticmethod
    def stack(name="stack", **kw):
        return _Stack(name)

    @staticmethod
    def hyperplane(input, output, name="hyperplane", bias_name=None, dot_product_name=None,  # noqa: A002
                   bias_weights=None, dot_product_weights=None):
        dn = dot_product_name or (name + "_w")
        bn = bias_name or (name + "_b")
        return _Component(lib.b200h_hyperplane_new(b(name), C.c_uint(input), C.c_uint(output), b(dn), b(bn),
                                                   b(dot_product_weights or dn), b(bias_weights or bn)))

    @staticmethod
    def dot_product(input, output, name="w", weights=None):  # noqa: A002
        return _Component(lib.b200h_dot_product_new(b(name), b(weights or name), C.c_uint(input), C.c_uint(output)))

    @staticmethod
    def bias(size, name="b", weights=None):
        return _Component(lib.b200h_bias_new(b(name), b(weights or name), C.c_uint(size)))

    @staticmethod
    def dropout(random, prob=0.5, value=0.0, norm=True, name="dropout", size=0):
        """ann.components.dropout{ name=, size=, prob=, value=, random=, norm= } (bind_ann_base.lua.cc:1643-1670).
        The component continues the stream of `random` from its current state (it keeps its own copy)."""
        return _Component(lib.b200h_dropout_new(b(name), random.h, C.c_float(prob), C.c_float(value), C.c_int(bool(norm)),
                                                C.c_uint(size)))

    @staticmethod
    def rewrap(size, name="rewrap"):
        return _Component(lib.b200h_rewrap_new(b(name), _ints(size), C.c_int(len(size))))

    @staticmethod
    def flatten(name="flatten"):
        return _Component(lib.b200h_flatten_new(b(name)))

    @staticmethod
    def convolution(kernel, n, name="conv", weights=None, step=None):
        return _Component(lib.b200h_convolution_new(b(name), b(weights or name), _ints(kernel),
                                                    _ints(step) if step else None, C.c_int(len(kernel)), C.c_int(n)))

    @staticmethod
    def convolution_bias(n, ndims=3, name="convb", weights=None):
        return _Component(lib.b200h_convolution_bias_new(b(name), b(weights or name), C.c_int(n)))

    @staticmethod
    def max_pooling(kernel, name="pool", step=None):
        return _Component(lib.b200h_max_pooling_new(b(name), _ints(kernel), _ints(step) if step else None,
                                                    C.c_int(len(kernel))))


class _AllAll:
    @staticmethod
    def generate(topology):
        return _Component(lib.b200h_mlp_generate(b(topology)))


class mlp:  # noqa: N801
    all_all = _AllAll


# ----------------------------------------------------------------------------- loss
class _Loss:
    def __init__(self, kind, size=0, TH=0.5):
        self.kind, self.size, self.TH = kind, size, TH


class loss:  # noqa: N801
    @staticmethod
    def mse(size=0):
        return _Loss(0, size)

    @staticmethod
    def cross_entropy(size=0):
        return _Loss(1, size)

    @staticmethod
    def multi_class_cross_entropy(size=0):
        if 0 < size < 3:
            raise B200Error(128, "Multi class cross entropy is only allowed for multi-class problems")
        return _Loss(2, size)


    @staticmethod
    def zero_one(size=0, TH=0.5):
        """ann.loss.zero_one(size, TH): 0/1 classification error, not differentiable
        (ann/loss/c_src/zero_one_loss_function.cc:39-132)."""
        return _Loss(3, size, TH)


# ----------------------------------------------------------------------------- optimizers
class _Optimizer:
    def __init__(self, name, **options):
        self.name, self.options = name, options


class optimizer:  # noqa: N801
    """ann.optimizer.*  (packages/ann/optimizer/lua_src/optimizer_{sgd,adagrad,rmsprop,adadelta}.lua)"""

    @staticmethod
    def sgd(**options):
        return _Optimizer("sgd", **options)

    @staticmethod
    def adagrad(**options):
        return _Optimizer("adagrad", **options)

    @staticmethod
    def rmsprop(**options):
        return _Optimizer("rmsprop", **options)

    @staticmethod
    def adadelta(**options):
        return _Optimizer("adadelta", **options)


# ----------------------------------------------------------------------------- trainer
class supervised_trainer:  # noqa: N801
    """trainable.supervised_trainer(ann_component, loss_function, bunch_size)
    -- packages/trainable/lua_src/supervised.lua:20-120."""

    def __init__(self, net, loss_function, bunch_size, optimizer=None, ctx=None):
        self.ctx = ctx or get_context()
        self.net = net
        self.bunch_size = bunch_size
        self.loss_function = loss_function
        self.h = nonnull(lib.b200h_trainer_new(self.ctx.h, net.h, C.c_int(loss_function.kind), C.c_int(bunch_size)))
        if loss_function.kind == 3:
            check(lib.b200h_trainer_set_loss_threshold(self.h, C.c_float(loss_function.TH)))
        self.is_built = False
        self.optimizer_name = "sgd"
        if optimizer is not None:
            self.set_optimizer(optimizer)

    def set_optimizer(self, optimizer):
        """optimizer: an ann.optimizer.* object (or its name).  Options are reset to that optimizer's
        defaults, its state is cleared."""
        name = optimizer if isinstance(optimizer, str) else optimizer.name
        check(lib.b200h_trainer_set_optimizer(self.h, b(name)))
        self.optimizer_name = name
        for k, v in ({} if isinstance(optimizer, str) else optimizer.options).items():
            self.set_option(k, v)

    def __del__(self):
        # (at interpreter shutdown the module globals may already be gone)
        if getattr(self, "h", None) and lib is not None:
            lib.b200h_trainer_free(self.h)
            self.h = None

    def build(self, input=0, output=0, **_):  # noqa: A002
        check(lib.b200h_trainer_build(self.h, C.c_uint(input or 0), C.c_uint(output or 0)))
        self.is_built = True
        return self

    def set_option(self, name, value):
        check(lib.b200h_trainer_set_option(self.h, b(name), C.c_double(value)))

    def get_option(self, name):
        v = C.c_double()
        check(lib.b200h_trainer_get_option(self.h, b(name), C.byref(v)))
        return v.value

    def set_layerwise_option(self, pattern, name, value):
        check(lib.b200h_trainer_set_layerwise_option(self.h, b(pattern), b(name), C.c_double(value)))

    def randomize_weights(self, random, inf, sup, use_fanin=False, use_fanout=False, name_match=None):  # noqa: A002
        check(lib.b200h_trainer_randomize_weights(self.h, random.h, C.c_double(inf), C.c_double(sup),
                                                  C.c_int(bool(use_fanin)), C.c_int(bool(use_fanout)), b(name_match)))

    def set_flag(self, flag, value):
        check(lib.b200h_trainer_set_flag(self.h, b(flag), C.c_int(int(value))))

    # -- weights ---------------------------------------------------------------------
    def weight_names(self):
        n = C.c_int()
        check(lib.b200h_trainer_num_weights(self.h, C.byref(n)))
        out = []
        buf = C.create_string_buffer(256)
        for i in range(n.value):
            check(lib.b200h_trainer_weight_name(self.h, C.c_int(i), buf, C.c_int(256)))
            out.append(buf.value.decode())
        return out

    def _dims(self, name):
        d = (C.c_int * 2)()
        check(lib.b200h_trainer_weight_dims(self.h, b(name), d))
        return d[0], d[1]

    def _get(self, name, which):
        a = np.empty(self._dims(name), dtype=_f32)
        check(lib.b200h_trainer_tensor_get(self.h, b(name), C.c_int(which), fptr(a)))
        return a

    def weights(self, name):
        return self._get(name, 0)

    def gradients(self, name):
        return self._get(name, 1)

    def updates(self, name):
        return self._get(name, 2)

    def set_weights(self, name, value, which=0):
        a = _as_f32(value)
        if a.shape != self._dims(name) and a.size != int(np.prod(self._dims(name))):
            raise B200Error(128, "Incorrect weights matrix dimensions")
        check(lib.b200h_trainer_tensor_set(self.h, b(name), C.c_int(which), fptr(a)))

    def num_parameters(self):
        n = C.c_uint64()
        check(lib.b200h_trainer_num_parameters(self.h, C.byref(n)))
        return n.value

    def get_input_size(self):
        n = C.c_int()
        check(lib.b200h_trainer_input_size(self.h, C.byref(n)))
        return n.value

    def get_output_size(self):
        n = C.c_int()
        check(lib.b200h_trainer_output_size(self.h, C.byref(n)))
        return n.value

    # -- steps -----------------------------------------------------------------------
    def _bunch(self, x, t):
        x, t = _as_f32(x), _as_f32(t)
        x = x.reshape(x.shape[0], -1)
        t = t.reshape(t.shape[0], -1)
        if x.shape[0] != t.shape[0]:
            raise B200Error(128, "Different token sizes found: input vs target")
        if x.shape[1] != self.get_input_size() or t.shape[1] != self.get_output_size():
            raise B200Error(128, "Incorrect patternSize: input %d (expected %d), target %d (expected %d)" % (
                x.shape[1], self.get_input_size(), t.shape[1], self.get_output_size()))
        return x, t

    def train_step(self, input, target, bunch_size=None, max_gradients_norm=None):  # noqa: A002
        """-> (bunch mean loss, per-pattern loss vector)  supervised.lua:725-821.
        bunch_size: the value of the gradient smoothing 1/sqrt(shared_count * bunch_size); as in the
        reference it defaults to the TRAINER's bunch_size, not to the row count of `input`
        (supervised.lua:757).  max_gradients_norm: global gradient-norm clip (supervised.lua:805-811)."""
        x, t = self._bunch(input, target)
        rows = np.empty(x.shape[0], dtype=_f32)
        l = C.c_float()
        check(lib.b200h_trainer_train_step_ex(self.h, fptr(x), fptr(t), C.c_int(x.shape[0]), C.c_int(int(bunch_size or 0)),
                                              C.c_double(float(max_gradients_norm or 0.0)), C.byref(l), fptr(rows)))
        return l.value, rows

    def validate_step(self, input, target):  # noqa: A002
        x, t = self._bunch(input, target)
        rows = np.empty(x.shape[0], dtype=_f32)
        l = C.c_float()
        check(lib.b200h_trainer_validate_step(self.h, fptr(x), fptr(t), C.c_int(x.shape[0]), C.byref(l), fptr(rows)))
        return l.value, rows

    def train_dataset(self, input_dataset, output_dataset, shuffle=None, **_):
        """-> loss:get_accum_loss() = (mean, variance)  supervised.lua:1149-1226.
        `shuffle` is a `random` object; its shuffle(n) draws the epoch order
        (trainable.lua:217-220)."""
        x, t = self._bunch(input_dataset, output_dataset)
        n = x.shape[0]
        order = shuffle.shuffle(n) if shuffle is not None else None
        m, v = C.c_float(), C.c_float()
        check(lib.b200h_trainer_train_dataset(
            self.h, fptr(x), fptr(t), C.c_int(n),
            order.ctypes.data_as(C.POINTER(C.c_int)) if order is not None else None, C.byref(m), C.byref(v)))
        return m.value, v.value

    def validate_dataset(self, input_dataset, output_dataset, **_):
        x, t = self._bunch(input_dataset, output_dataset)
        m, v = C.c_float(), C.c_float()
        check(lib.b200h_trainer_validate_dataset(self.h, fptr(x), fptr(t), C.c_int(x.shape[0]), C.byref(m), C.byref(v)))
        return m.value, v.value

    def calculate(self, input):  # noqa: A002
        x = _as_f32(input)
        x = x.reshape(x.shape[0], -1)
        y = np.empty((x.shape[0], self.get_output_size()), dtype=_f32)
        check(lib.b200h_trainer_calculate(self.h, fptr(x), C.c_int(x.shape[0]), fptr(y)))
        return y

    def use_dataset(self, input_dataset):
        """forward only over a dataset, bunch by bunch -> [n, output_size]  (supervised.lua:1291-1430)"""
        x = _as_f32(input_dataset)
        x = x.reshape(x.shape[0], -1)
        if x.shape[1] != self.get_input_size():
            raise B200Error(128, "Incorrect patternSize: input %d (expected %d)" % (x.shape[1], self.get_input_size()))
        y = np.empty((x.shape[0], self.get_output_size()), dtype=_f32)
        check(lib.b200h_trainer_use_dataset(self.h, fptr(x), C.c_int(x.shape[0]), fptr(y)))
        return y

    # -- checkpoint / resume (optimizer_sgd.lua:102-119: options + count + update; supervised.lua:349-395) ------
    def get_count(self):
        c = C.c_int64()
        check(lib.b200h_trainer_get_count(self.h, C.byref(c)))
        return c.value

    def set_count(self, count):
        check(lib.b200h_trainer_set_count(self.h, C.c_int64(int(count))))

    def optimizer_state(self, name, which):
        """which: 2 = update (sgd momentum buffer / rmsprop Eupdates / adadelta update), 3 = Egradients / Erms,
        4 = adadelta Eupdates"""
        return self._get(name, which)

    def state_dict(self):
        """Everything a resumed run needs, as host arrays: weights, optimizer state tensors, count, options."""
        kinds = {"sgd": (2,), "adagrad": (3,), "rmsprop": (2, 3), "adadelta": (2, 3, 4)}[self.optimizer_name]
        return {"optimizer": self.optimizer_name, "count": self.get_count(),
                "weights": {n: self.weights(n) for n in self.weight_names()},
                "state": {k: {n: self._get(n, k) for n in self.weight_names()} for k in kinds}}

    def load_state_dict(self, sd):
        if sd["optimizer"] != self.optimizer_name:
            raise B200Error(128, "checkpoint of optimizer %s loaded into %s" % (sd["optimizer"], self.optimizer_name))
        for n, w in sd["weights"].items():
            self.set_weights(n, w)
        for k, d in sd["state"].items():
            for n, v in d.items():
                self.set_weights(n, v, which=int(k))
        self.set_count(sd["count"])

    def component_token(self, component, which="output"):
        """get_input/get_output/get_error_input/get_error_output of a named component after
        the last step (bind_ann_base.lua.cc:392-565)."""
        w = {"input": 0, "output": 1, "error_input": 2, "error_output": 3}[which]
        dims = (C.c_int * 4)()
        nd = C.c_int()
        check(lib.b200h_trainer_component_token(self.h, b(component), C.c_int(w), None, dims, C.byref(nd)))
        a = np.empty(tuple(dims[i] for i in range(nd.value)), dtype=_f32)
        check(lib.b200h_trainer_component_token(self.h, b(component), C.c_int(w), fptr(a), dims, C.byref(nd)))
        return a

    def norm2(self, pattern=".*"):
        import re
        best = 0.0
        for n in self.weight_names():
            if re.search(pattern.replace("%", "\\"), n):
                w = self.weights(n).astype(np.float64)
                best = max(best, float(np.sqrt((w * w).sum(axis=1)).max()))
        return best

    # -- pipelined / data parallel -------------------------------------------------------
    def stage(self, x_pinned, t_pinned, bunch):
        check(lib.b200h_trainer_stage(self.h, fptr(x_pinned), fptr(t_pinned), C.c_int(bunch)))

    def step_staged(self, bunch):
        check(lib.b200h_trainer_step_staged(self.h, C.c_int(bunch)))

    def loss_reset(self):
        check(lib.b200h_trainer_loss_reset(self.h))

    def loss_get(self):
        m, v = C.c_float(), C.c_float()
        check(lib.b200h_trainer_loss_get(self.h, C.byref(m), C.byref(v)))
        return m.value, v.value

    def set_data_parallel(self, nranks, rank):
        check(lib.b200h_trainer_set_data_parallel(self.h, C.c_int(nranks), C.c_int(rank)))

    def broadcast_weights(self):
        check(lib.b200h_trainer_broadcast_weights(self.h))


class trainable:  # noqa: N801
    supervised_trainer = supervised_trainer
