"""TEST INFRASTRUCTURE: ctypes face of oracle/_ref/libaprilref.so.

libaprilref.so is the reference's own C++ (components, loss functions, matrix,
BLAS wrappers) compiled in place from /root/reference by oracle/ref_build/
Makefile.  This module only marshals numpy arrays in and out of it, so that
tests/test_oracle_vs_reference.py can put the reference's numbers beside
oracle/april.py's.  Only tests/ may import it; the product never does.

The class and method names follow the reference's Lua API
(packages/ann/ann/binding/bind_ann_base.lua.cc:287-2187,
packages/ann/loss/binding/bind_loss_functions.lua.cc:60-140).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# APRILREF_LIB selects another build of the same C face: the integration builds
# (integration/_build/libaprilref_{shim,b200}.so) are driven through this module too.
LIB_PATH = os.environ.get("APRILREF_LIB") or os.path.join(HERE, "_ref", "libaprilref.so")

_lib = None
_F = C.POINTER(C.c_float)
_I = C.POINTER(C.c_int)


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        for name in ("ref_stack_new", "ref_hyperplane_new", "ref_dot_product_new", "ref_bias_new",
                     "ref_actf_new", "ref_prelu_new", "ref_convolution_new",
                     "ref_convolution_bias_new", "ref_max_pooling_new", "ref_flatten_new",
                     "ref_rewrap_new", "ref_dropout_new", "ref_net_build", "ref_loss_new",
                     "ref_random_new"):
            getattr(L, name).restype = C.c_void_p
        L.ref_last_error.restype = C.c_char_p
        L.ref_random_rand.restype = C.c_double
        L.ref_random_randint.restype = C.c_uint
        L.ref_stack_push.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_hyperplane_new.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_int]
        L.ref_dot_product_new.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_int]
        L.ref_bias_new.argtypes = [C.c_int, C.c_char_p]
        L.ref_actf_new.argtypes = [C.c_char_p, C.c_float, C.c_float]
        L.ref_prelu_new.argtypes = [C.c_int, C.c_int, C.c_char_p]
        L.ref_convolution_new.argtypes = [C.c_int, _I, _I, C.c_int, C.c_char_p]
        L.ref_convolution_bias_new.argtypes = [C.c_int, C.c_int, C.c_char_p]
        L.ref_max_pooling_new.argtypes = [C.c_int, _I, _I]
        L.ref_rewrap_new.argtypes = [_I, C.c_int]
        L.ref_dropout_new.argtypes = [C.c_uint, C.c_float, C.c_float, C.c_int]
        L.ref_component_free.argtypes = [C.c_void_p]
        L.ref_component_set_use_cuda.argtypes = [C.c_void_p, C.c_int]
        L.ref_set_use_cuda_default.argtypes = [C.c_int]
        L.ref_net_build.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ref_net_free.argtypes = [C.c_void_p]
        L.ref_net_tensor_get.argtypes = [C.c_void_p, C.c_int, C.c_char_p, _F, C.c_int, _I, _I]
        L.ref_net_weight_set.argtypes = [C.c_void_p, C.c_char_p, _F, C.c_int]
        L.ref_net_shared_count.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_net_forward.argtypes = [C.c_void_p, _F, C.c_int, _I, C.c_int, _F, C.c_int, _I, _I]
        L.ref_net_backprop.argtypes = [C.c_void_p, _F, C.c_int, _I, _F, C.c_int, _I, _I]
        L.ref_net_compute_gradients.argtypes = [C.c_void_p]
        L.ref_net_reset.argtypes = [C.c_void_p, C.c_uint]
        L.ref_loss_new.argtypes = [C.c_char_p, C.c_int, C.c_float]
        L.ref_loss_free.argtypes = [C.c_void_p]
        L.ref_loss_compute.argtypes = [C.c_void_p, _F, _F, C.c_int, C.c_int, C.c_int, _F]
        L.ref_loss_gradient.argtypes = [C.c_void_p, _F, _F, C.c_int, C.c_int, _F]
        L.ref_random_new.argtypes = [C.c_uint]
        L.ref_random_free.argtypes = [C.c_void_p]
        L.ref_random_rand.argtypes = [C.c_void_p]
        L.ref_random_randint.argtypes = [C.c_void_p]
        L.ref_random_rand_n.restype = C.c_double
        L.ref_random_rand_n.argtypes = [C.c_void_p, C.c_double]
        L.ref_random_randint_n.restype = C.c_uint
        L.ref_random_randint_n.argtypes = [C.c_void_p, C.c_uint]
        L.ref_random_shuffle.argtypes = [C.c_void_p, C.c_int, _I]
        L.ref_gemm.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _F, _F,
                               C.c_float, _F]
        if L.ref_init() != 0:
            raise RuntimeError("reference library failed to initialise")
        _lib = L
    return _lib


REF_FAILED = -2 ** 31


class ReferenceError_(RuntimeError):
    """An ERROR_EXIT raised inside the reference (util/c_src/error_print.h:55-77)."""


def _raise():
    raise ReferenceError_(lib().ref_last_error().decode(errors="replace").strip())


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _fp(a):
    return a.ctypes.data_as(_F)


def _ints(v):
    return (C.c_int * len(v))(*[int(i) for i in v])


def _b(s):
    return s.encode() if s is not None else None


# ---- component constructors (ann.components.*) ---------------------------

def built_with_cuda():
    return bool(lib().ref_built_with_cuda())


def set_use_cuda_default(flag):
    """mathcore.set_use_cuda_default(flag)."""
    lib().ref_set_use_cuda_default(int(flag))


def set_use_cuda(component, flag):
    """component:set_use_cuda(flag); call before Net(...) builds it."""
    lib().ref_component_set_use_cuda(component, int(flag))
    return component


def stack():
    return lib().ref_stack_new()


def push(s, *comps):
    """The stack takes its own reference on each component; the constructor's is released here."""
    for c in comps:
        lib().ref_stack_push(s, c)
        lib().ref_component_free(c)
    return s


def hyperplane(input, output, dot_product_weights, bias_weights, transpose=False):
    return lib().ref_hyperplane_new(input, output, _b(dot_product_weights), _b(bias_weights), int(transpose))


def dot_product(input, output, weights, transpose=False):
    return lib().ref_dot_product_new(input, output, _b(weights), int(transpose))


def b200_dot_product(input, output, weights, transpose=False):
    """integration builds only: the dot-product component whose dense methods call libb200ann.so
    (integration/ann/b200_dot_product_component.h)."""
    fn = lib().ref_b200_dot_product_new
    fn.restype = C.c_void_p
    fn.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_int]
    return fn(input, output, _b(weights), int(transpose))


def bias(size, weights):
    return lib().ref_bias_new(size, _b(weights))


def actf(kind, p0=0.0, p1=0.0):
    h = lib().ref_actf_new(_b(kind), p0, p1)
    if not h:
        raise ValueError("unknown activation " + kind)
    return h


def prelu(size, weights, scalar=False):
    return lib().ref_prelu_new(int(scalar), size, _b(weights))


def convolution(kernel, n, weights, step=None):
    step = step or [1] * len(kernel)
    return lib().ref_convolution_new(len(kernel), _ints(kernel), _ints(step), n, _b(weights))


def convolution_bias(ndims, n, weights):
    return lib().ref_convolution_bias_new(ndims, n, _b(weights))


def max_pooling(kernel, step=None):
    step = step or kernel
    return lib().ref_max_pooling_new(len(kernel), _ints(kernel), _ints(step))


def flatten():
    return lib().ref_flatten_new()


def rewrap(dims):
    return lib().ref_rewrap_new(_ints(dims), len(dims))


def dropout(seed, prob=0.5, value=0.0, norm=True):
    return lib().ref_dropout_new(seed, value, prob, int(norm))


class Net:
    """A built component tree with its weights and gradient dictionaries."""

    def __init__(self, root, input_size=0, output_size=0):
        self.h = lib().ref_net_build(root, input_size, output_size)
        if not self.h:
            _raise()
        lib().ref_component_free(root)      # the net holds its own reference on the root

    def close(self):
        if self.h:
            lib().ref_net_free(self.h)
            self.h = None

    def _tensor(self, which, name):
        nd = C.c_int(0)
        dims = (C.c_int * 8)()
        n = lib().ref_net_tensor_get(self.h, which, _b(name), None, 0, C.byref(nd), dims)
        if n == 0:
            return None
        out = np.empty(n, np.float32)
        lib().ref_net_tensor_get(self.h, which, _b(name), _fp(out), n, C.byref(nd), dims)
        return out.reshape([dims[i] for i in range(nd.value)])

    def weight(self, name):
        return self._tensor(0, name)

    def gradient(self, name):
        return self._tensor(1, name)

    def set_weight(self, name, value):
        v = _f32(value).ravel()
        if lib().ref_net_weight_set(self.h, _b(name), _fp(v), v.size) != 0:
            raise ValueError("no weight %r of %d elements" % (name, v.size))

    def shared_count(self, name):
        return lib().ref_net_shared_count(self.h, _b(name))

    def _run(self, fn, x, *extra, out_elems=None):
        x = _f32(x)
        nd = C.c_int(0)
        dims = (C.c_int * 8)()
        cap = int(out_elems) if out_elems else 1 << 24
        out = np.empty(cap, np.float32)
        n = fn(self.h, _fp(x), x.ndim, _ints(x.shape), *extra, _fp(out), cap, C.byref(nd), dims)
        if n == REF_FAILED:
            _raise()
        if n < 0:
            raise RuntimeError("output of %d elements exceeds the marshalling buffer" % -n)
        if n == 0:
            return None
        return out[:n].reshape([dims[i] for i in range(nd.value)]).copy()

    def forward(self, x, during_training=False, out_elems=None):
        """out_elems: size of the marshalling buffer when the output exceeds 16 Mi floats"""
        return self._run(lib().ref_net_forward, x, int(during_training), out_elems=out_elems)

    def backprop(self, e, out_elems=None):
        return self._run(lib().ref_net_backprop, e, out_elems=out_elems)

    def compute_gradients(self):
        if lib().ref_net_compute_gradients(self.h) == REF_FAILED:
            _raise()

    def reset(self, it=0):
        lib().ref_net_reset(self.h, it)


class Loss:
    """ann.loss.{mse,multi_class_cross_entropy,cross_entropy,zero_one}."""

    def __init__(self, kind, size=0, param=0.5):
        self.h = lib().ref_loss_new(_b(kind), size, param)
        if not self.h:
            raise ValueError("unknown loss " + kind)

    def loss_rows(self, out, tgt):
        out, tgt = _f32(out), _f32(tgt)
        if tgt.ndim == 1:
            tgt = tgt[:, None]
        v = np.empty(out.shape[0], np.float32)
        rc = lib().ref_loss_compute(self.h, _fp(out), _fp(tgt), out.shape[0], out.shape[1],
                                    tgt.shape[1], _fp(v))
        if rc == REF_FAILED:
            _raise()
        if rc != 0:
            raise RuntimeError("compute_loss failed (%d)" % rc)
        return v

    def gradient(self, out, tgt):
        out, tgt = _f32(out), _f32(tgt)
        g = np.empty_like(out)
        rc = lib().ref_loss_gradient(self.h, _fp(out), _fp(tgt), out.shape[0], out.shape[1], _fp(g))
        if rc == REF_FAILED:
            _raise()
        if rc != 0:
            raise RuntimeError("gradient failed")
        return g


class Random:
    """The reference's MTRand (packages/basics/random/c_src/MersenneTwister.h)."""

    def __init__(self, seed):
        self.h = lib().ref_random_new(seed)

    def rand(self):
        return lib().ref_random_rand(self.h)

    def randint(self, n=None):
        return lib().ref_random_randint(self.h) if n is None else lib().ref_random_randint_n(self.h, n)

    def rand_n(self, n):
        return lib().ref_random_rand_n(self.h, n)

    def shuffle(self, size):
        out = (C.c_int * size)()
        lib().ref_random_shuffle(self.h, size, out)
        return list(out)


def gemm(ta, tb, alpha, a, b, beta, c):
    """C = alpha op(A) op(B) + beta C via MatrixExt::BLAS::matGemm."""
    a, b = _f32(a), _f32(b)
    c = _f32(c).copy()
    m, n = c.shape
    k = a.shape[0] if ta else a.shape[1]
    if lib().ref_gemm(int(ta), int(tb), m, n, k, alpha, _fp(a), _fp(b), beta, _fp(c)) == REF_FAILED:
        _raise()
    return c
