"""Kernel-level entry points of the C ABI (include/b200ann.h) for numpy callers.

Each function uploads its numpy operands, runs ONE C-ABI call on the device and returns the
result as numpy -- the shape in which the reference's own unit tests exercise
matrix:gemm / ann.components / ann.loss (packages/basics/matrix/test/test_gemm.lua,
packages/ann/loss/test/test.lua).  Used by the parity tests; training goes through
`supervised_trainer`, which keeps everything device-resident.
"""
import ctypes as C

import numpy as np

from . import get_context
from ._lib import lib, check

_f32 = np.float32
ACT = {None: 0, "none": 0, "logistic": 1, "tanh": 2, "relu": 3, "softmax": 4, "log_softmax": 5, "linear": 6}


class DeviceArray:
    """A float32 (or int32/float64) device buffer from the context's caching pool."""

    def __init__(self, ctx, shape, dtype=_f32):
        self.ctx = ctx
        self.shape = tuple(int(s) for s in np.atleast_1d(shape))
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p = C.c_void_p()
        check(lib.b200_malloc(ctx.h, C.byref(p), C.c_size_t(max(self.nbytes, 4))))
        self.ptr = p

    @classmethod
    def from_numpy(cls, ctx, a, dtype=_f32):
        a = np.ascontiguousarray(a, dtype=dtype)
        d = cls(ctx, a.shape, dtype)
        check(lib.b200_memcpy_h2d(ctx.h, d.ptr, a.ctypes.data_as(C.c_void_p), C.c_size_t(a.nbytes)))
        check(lib.b200_sync(ctx.h))  # `a` may be a temporary
        return d

    def numpy(self):
        out = np.empty(self.shape, dtype=self.dtype)
        check(lib.b200_memcpy_d2h(self.ctx.h, out.ctypes.data_as(C.c_void_p), self.ptr, C.c_size_t(self.nbytes)))
        check(lib.b200_sync(self.ctx.h))
        return out

    def zero(self):
        check(lib.b200_memset_zero(self.ctx.h, self.ptr, C.c_size_t(self.nbytes)))

    def free(self):
        if self.ptr:
            lib.b200_free(self.ctx.h, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _dev(ctx, a):
    return DeviceArray.from_numpy(ctx, a)


def sgemm(transA, transB, alpha, A, B, beta=0.0, Cin=None, ctx=None):
    """C = alpha*op(A)*op(B) + beta*C, row-major (matrix:gemm{...},
    packages/basics/matrix/binding/matrix_binding.h:1703)."""
    ctx = ctx or get_context()
    A, B = np.ascontiguousarray(A, _f32), np.ascontiguousarray(B, _f32)
    M, K = (A.shape[1], A.shape[0]) if transA else A.shape
    K2, N = (B.shape[1], B.shape[0]) if transB else B.shape
    if K != K2:
        raise ValueError("Incorrect matrix sizes")
    dA, dB = _dev(ctx, A), _dev(ctx, B)
    dC = _dev(ctx, Cin) if Cin is not None else DeviceArray(ctx, (M, N))
    if Cin is None:
        dC.zero()
    check(lib.b200_sgemm(ctx.h, C.c_int(int(transA)), C.c_int(int(transB)), C.c_int(M), C.c_int(N), C.c_int(K),
                         C.c_float(alpha), dA.ptr, C.c_int(A.shape[1]), dB.ptr, C.c_int(B.shape[1]),
                         C.c_float(beta), dC.ptr, C.c_int(N)))
    return dC.numpy()


def linear_fwd(X, W, bias=None, act=None, ctx=None):
    ctx = ctx or get_context()
    X, W = np.ascontiguousarray(X, _f32), np.ascontiguousarray(W, _f32)
    M, K = X.shape
    N = W.shape[0]
    dX, dW = _dev(ctx, X), _dev(ctx, W)
    db = _dev(ctx, np.asarray(bias, _f32).reshape(-1)) if bias is not None else None
    dY = DeviceArray(ctx, (M, N))
    check(lib.b200_linear_fwd(ctx.h, C.c_int(M), C.c_int(N), C.c_int(K), dX.ptr, C.c_int(K), dW.ptr, C.c_int(K),
                              db.ptr if db else None, C.c_int(ACT[act]), dY.ptr, C.c_int(N)))
    return dY.numpy()


def linear_bwd_data(dY, W, act_prev=None, Yprev=None, ctx=None):
    ctx = ctx or get_context()
    dY, W = np.ascontiguousarray(dY, _f32), np.ascontiguousarray(W, _f32)
    M, N = dY.shape
    K = W.shape[1]
    ddY, dW = _dev(ctx, dY), _dev(ctx, W)
    dYp = _dev(ctx, Yprev) if Yprev is not None else None
    dX = DeviceArray(ctx, (M, K))
    check(lib.b200_linear_bwd_data(ctx.h, C.c_int(M), C.c_int(N), C.c_int(K), ddY.ptr, C.c_int(N), dW.ptr, C.c_int(K),
                                   C.c_int(ACT[act_prev]), dYp.ptr if dYp else None, C.c_int(K), dX.ptr, C.c_int(K)))
    return dX.numpy()


def linear_bwd_weight(dY, X, scale=1.0, beta=0.0, dW0=None, db0=None, want_db=True, ctx=None):
    ctx = ctx or get_context()
    dY, X = np.ascontiguousarray(dY, _f32), np.ascontiguousarray(X, _f32)
    M, N = dY.shape
    K = X.shape[1]
    ddY, dX = _dev(ctx, dY), _dev(ctx, X)
    dW = _dev(ctx, dW0) if dW0 is not None else DeviceArray(ctx, (N, K))
    db = (_dev(ctx, db0) if db0 is not None else DeviceArray(ctx, (N,))) if want_db else None
    if dW0 is None:
        dW.zero()
    if db is not None and db0 is None:
        db.zero()
    check(lib.b200_linear_bwd_weight(ctx.h, C.c_int(M), C.c_int(N), C.c_int(K), ddY.ptr, C.c_int(N), dX.ptr, C.c_int(K),
                                     C.c_float(scale), C.c_float(beta), dW.ptr, C.c_int(K), db.ptr if db else None))
    return dW.numpy(), (db.numpy() if db is not None else None)


def actf_fwd(act, x, ctx=None):
    ctx = ctx or get_context()
    x = np.ascontiguousarray(x, _f32)
    dx = _dev(ctx, x)
    dy = DeviceArray(ctx, x.shape)
    if act in ("softmax", "log_softmax"):
        f = lib.b200_softmax_fwd if act == "softmax" else lib.b200_log_softmax_fwd
        x2 = x.reshape(x.shape[0], -1)
        check(f(ctx.h, C.c_int(x2.shape[0]), C.c_int(x2.shape[1]), dx.ptr, dy.ptr))
    else:
        check(lib.b200_actf_fwd(ctx.h, C.c_int(ACT[act]), C.c_size_t(x.size), dx.ptr, dy.ptr))
    return dy.numpy()


def actf_bwd(act, y, dy, ctx=None):
    ctx = ctx or get_context()
    y, dy = np.ascontiguousarray(y, _f32), np.ascontiguousarray(dy, _f32)
    d_y, d_dy = _dev(ctx, y), _dev(ctx, dy)
    d_dx = DeviceArray(ctx, y.shape)
    if act == "softmax":
        check(lib.b200_softmax_bwd(ctx.h, C.c_int(y.shape[0]), C.c_int(y.shape[1]), d_y.ptr, d_dy.ptr, d_dx.ptr))
    else:
        check(lib.b200_actf_bwd(ctx.h, C.c_int(ACT[act]), C.c_size_t(y.size), d_y.ptr, d_dy.ptr, d_dx.ptr))
    return d_dx.numpy()


def loss_and_grad(kind, out, target, ctx=None):
    """kind in {'mse','cross_entropy','multi_class_cross_entropy'} -> (loss rows, gradient)."""
    ctx = ctx or get_context()
    out, target = np.ascontiguousarray(out, _f32), np.ascontiguousarray(target, _f32)
    M, Cc = out.shape
    d_o, d_t = _dev(ctx, out), _dev(ctx, target)
    rows, grad = DeviceArray(ctx, (M,)), DeviceArray(ctx, (M, Cc))
    f = {"mse": lib.b200_mse_loss_grad, "cross_entropy": lib.b200_ce_loss_grad,
         "multi_class_cross_entropy": lib.b200_mcce_loss_grad}[kind]
    check(f(ctx.h, C.c_int(M), C.c_int(Cc), d_o.ptr, d_t.ptr, rows.ptr, grad.ptr))
    return rows.numpy(), grad.numpy()


def log_softmax_mcce_fused(logits, target, ctx=None):
    ctx = ctx or get_context()
    logits, target = np.ascontiguousarray(logits, _f32), np.ascontiguousarray(target, _f32)
    M, Cc = logits.shape
    d_z, d_t = _dev(ctx, logits), _dev(ctx, target)
    logp, rows, grad = DeviceArray(ctx, (M, Cc)), DeviceArray(ctx, (M,)), DeviceArray(ctx, (M, Cc))
    check(lib.b200_log_softmax_mcce_fused(ctx.h, C.c_int(M), C.c_int(Cc), d_z.ptr, d_t.ptr, logp.ptr, rows.ptr, grad.ptr))
    return logp.numpy(), rows.numpy(), grad.numpy()


def output_layer_fused(X, W, bias, target, act_prev=None, want_dx=True, ctx=None):
    """dot_product + bias + log_softmax + MCCE rows + gradient + data gradient of the layer below in one
    launch -> (logits, logp, loss rows, grad, dX or None)."""
    ctx = ctx or get_context()
    X, W, target = (np.ascontiguousarray(a, _f32) for a in (X, W, target))
    M, K = X.shape
    N = W.shape[0]
    d_x, d_w, d_t = _dev(ctx, X), _dev(ctx, W), _dev(ctx, target)
    d_b = _dev(ctx, np.ascontiguousarray(bias, _f32).reshape(-1)) if bias is not None else None
    logits, logp, rows, grad = (DeviceArray(ctx, (M, N)), DeviceArray(ctx, (M, N)), DeviceArray(ctx, (M,)),
                                DeviceArray(ctx, (M, N)))
    dx = DeviceArray(ctx, (M, K)) if want_dx else None
    check(lib.b200_output_layer_fused(ctx.h, C.c_int(M), C.c_int(N), C.c_int(K), d_x.ptr, C.c_int(K), d_w.ptr, C.c_int(K),
                                      d_b.ptr if d_b is not None else None, d_t.ptr, logits.ptr, logp.ptr, rows.ptr,
                                      grad.ptr, C.c_int(ACT[act_prev] if act_prev else 0), dx.ptr if dx is not None else None,
                                      C.c_int(K)))
    return logits.numpy(), logp.numpy(), rows.numpy(), grad.numpy(), (dx.numpy() if dx is not None else None)


def conv2d_fwd(x, w, kernel, step=(1, 1), bias=None, act=None, ctx=None):
    ctx = ctx or get_context()
    x, w = np.ascontiguousarray(x, _f32), np.ascontiguousarray(w, _f32)
    B, Cc, H, W = x.shape
    n = w.shape[0]
    kh, kw = kernel
    sh, sw = step
    oH, oW = (H - kh) // sh + 1, (W - kw) // sw + 1
    dx, dw = _dev(ctx, x), _dev(ctx, w)
    db = _dev(ctx, np.asarray(bias, _f32).reshape(-1)) if bias is not None else None
    dy = DeviceArray(ctx, (B, n, oH, oW))
    check(lib.b200_conv2d_fwd(ctx.h, *[C.c_int(v) for v in (B, Cc, H, W, n, kh, kw, sh, sw)], dx.ptr, dw.ptr,
                              db.ptr if db else None, C.c_int(ACT[act]), dy.ptr))
    return dy.numpy()


def conv2d_bwd_data(dy, w, x_shape, kernel, step=(1, 1), ctx=None):
    ctx = ctx or get_context()
    dy, w = np.ascontiguousarray(dy, _f32), np.ascontiguousarray(w, _f32)
    B, Cc, H, W = x_shape
    n = w.shape[0]
    kh, kw = kernel
    sh, sw = step
    ddy, dw = _dev(ctx, dy), _dev(ctx, w)
    dx = DeviceArray(ctx, x_shape)
    check(lib.b200_conv2d_bwd_data(ctx.h, *[C.c_int(v) for v in (B, Cc, H, W, n, kh, kw, sh, sw)], ddy.ptr, dw.ptr,
                                   dx.ptr))
    return dx.numpy()


def conv2d_bwd_weight(dy, x, kernel, step=(1, 1), scale=1.0, ctx=None):
    ctx = ctx or get_context()
    dy, x = np.ascontiguousarray(dy, _f32), np.ascontiguousarray(x, _f32)
    B, Cc, H, W = x.shape
    n = dy.shape[1]
    kh, kw = kernel
    sh, sw = step
    ddy, dx = _dev(ctx, dy), _dev(ctx, x)
    dw, db = DeviceArray(ctx, (n, Cc * kh * kw)), DeviceArray(ctx, (n,))
    check(lib.b200_conv2d_bwd_weight(ctx.h, *[C.c_int(v) for v in (B, Cc, H, W, n, kh, kw, sh, sw)], ddy.ptr, dx.ptr,
                                     C.c_float(scale), C.c_float(0.0), dw.ptr, db.ptr))
    return dw.numpy(), db.numpy()


def maxpool_fwd(x, kernel, step=None, ctx=None):
    ctx = ctx or get_context()
    x = np.ascontiguousarray(x, _f32)
    B, Cc, H, W = x.shape
    kh, kw = kernel
    sh, sw = step or kernel
    oH, oW = (H - kh) // sh + 1, (W - kw) // sw + 1
    dx = _dev(ctx, x)
    dy = DeviceArray(ctx, (B, Cc, oH, oW))
    arg = DeviceArray(ctx, (B, Cc, oH, oW), np.int32)
    check(lib.b200_maxpool_fwd(ctx.h, *[C.c_int(v) for v in (B, Cc, H, W, kh, kw, sh, sw)], dx.ptr, dy.ptr, arg.ptr))
    return dy.numpy(), arg.numpy()


def maxpool_bwd(dy, argmax, x_shape, kernel, step=None, ctx=None):
    ctx = ctx or get_context()
    dy = np.ascontiguousarray(dy, _f32)
    B, Cc, H, W = x_shape
    kh, kw = kernel
    sh, sw = step or kernel
    ddy = _dev(ctx, dy)
    darg = DeviceArray.from_numpy(ctx, argmax, np.int32)
    dx = DeviceArray(ctx, x_shape)
    check(lib.b200_maxpool_bwd(ctx.h, *[C.c_int(v) for v in (B, Cc, H, W, kh, kw, sh, sw)], ddy.ptr, darg.ptr, dx.ptr))
    return dx.numpy()


# ---- BLAS level 1/2 seam (mathcore/c_src/cblas_headers.h:240-535) ----------------------------------
def sgemv(transA, alpha, A, x, beta=0.0, y0=None, incx=1, incy=1, ctx=None):
    """y = alpha*op(A)*x + beta*y, A [M,N] row-major (matrix:gemv, gemv.cu:44).  x / y are passed with
    their BLAS increments: element i of the vector lives at index i*inc of the buffer."""
    ctx = ctx or get_context()
    A = np.ascontiguousarray(A, _f32)
    M, N = A.shape
    rows = N if transA else M
    xb = np.ascontiguousarray(x, _f32).reshape(-1)
    yb = np.zeros(rows * incy, _f32) if y0 is None else np.ascontiguousarray(y0, _f32).reshape(-1).copy()
    dA, dx, dy = _dev(ctx, A), _dev(ctx, xb), _dev(ctx, yb)
    check(lib.b200_sgemv(ctx.h, C.c_int(int(transA)), C.c_int(M), C.c_int(N), C.c_float(alpha), dA.ptr, C.c_int(N),
                         dx.ptr, C.c_int(incx), C.c_float(beta), dy.ptr, C.c_int(incy)))
    return dy.numpy()


def sger(alpha, x, y, A0, incx=1, incy=1, ctx=None):
    """A += alpha * x y^T (matrix:ger, ger.cu:42)."""
    ctx = ctx or get_context()
    A0 = np.ascontiguousarray(A0, _f32)
    M, N = A0.shape
    dx, dy = _dev(ctx, np.ascontiguousarray(x, _f32).reshape(-1)), _dev(ctx, np.ascontiguousarray(y, _f32).reshape(-1))
    dA = _dev(ctx, A0)
    check(lib.b200_sger(ctx.h, C.c_int(M), C.c_int(N), C.c_float(alpha), dx.ptr, C.c_int(incx), dy.ptr, C.c_int(incy),
                        dA.ptr, C.c_int(N)))
    return dA.numpy()


def saxpy(alpha, x, y, ctx=None):
    ctx = ctx or get_context()
    x, y = np.ascontiguousarray(x, _f32), np.ascontiguousarray(y, _f32)
    dx, dy = _dev(ctx, x), _dev(ctx, y)
    check(lib.b200_saxpy(ctx.h, C.c_size_t(x.size), C.c_float(alpha), dx.ptr, dy.ptr))
    return dy.numpy()


def sscal(alpha, x, ctx=None):
    ctx = ctx or get_context()
    x = np.ascontiguousarray(x, _f32)
    dx = _dev(ctx, x)
    check(lib.b200_sscal(ctx.h, C.c_size_t(x.size), C.c_float(alpha), dx.ptr))
    return dx.numpy()


def scopy(x, ctx=None):
    ctx = ctx or get_context()
    x = np.ascontiguousarray(x, _f32)
    dx, dy = _dev(ctx, x), DeviceArray(ctx, x.shape)
    check(lib.b200_scopy(ctx.h, C.c_size_t(x.size), dx.ptr, dy.ptr))
    return dy.numpy()


def cmul(x, y, ctx=None):
    """y *= x (matCmul)."""
    ctx = ctx or get_context()
    x, y = np.ascontiguousarray(x, _f32), np.ascontiguousarray(y, _f32)
    dx, dy = _dev(ctx, x), _dev(ctx, y)
    check(lib.b200_cmul(ctx.h, C.c_size_t(x.size), dx.ptr, dy.ptr))
    return dy.numpy()


def ssum(x, ctx=None):
    ctx = ctx or get_context()
    x = np.ascontiguousarray(x, _f32)
    dx, out = _dev(ctx, x), DeviceArray(ctx, (1,))
    check(lib.b200_sum(ctx.h, C.c_size_t(x.size), dx.ptr, out.ptr))
    return float(out.numpy()[0])


def nrm2sq(xs, ctx=None):
    """sum over the given arrays of sum(x^2), accumulated in one device scalar (the global gradient
    norm of supervised.lua:805-811 is its square root)."""
    ctx = ctx or get_context()
    out = DeviceArray(ctx, (1,))
    out.zero()
    keep = []
    for x in xs:
        x = np.ascontiguousarray(x, _f32)
        dx = _dev(ctx, x)
        keep.append(dx)
        check(lib.b200_nrm2sq(ctx.h, C.c_size_t(x.size), dx.ptr, out.ptr))
    return float(out.numpy()[0])


def bias_fwd(x, b, ctx=None):
    ctx = ctx or get_context()
    x = np.ascontiguousarray(x, _f32)
    dx, db, dy = _dev(ctx, x), _dev(ctx, np.ascontiguousarray(b, _f32).reshape(-1)), DeviceArray(ctx, x.shape)
    check(lib.b200_bias_fwd(ctx.h, C.c_int(x.shape[0]), C.c_int(x.shape[1]), dx.ptr, db.ptr, dy.ptr))
    return dy.numpy()


def conv_bias_fwd(x, b, ctx=None):
    ctx = ctx or get_context()
    x = np.ascontiguousarray(x, _f32)
    B, n, H, W = x.shape
    dx, db, dy = _dev(ctx, x), _dev(ctx, np.ascontiguousarray(b, _f32).reshape(-1)), DeviceArray(ctx, x.shape)
    check(lib.b200_conv_bias_fwd(ctx.h, C.c_int(B), C.c_int(n), C.c_int(H * W), dx.ptr, db.ptr, dy.ptr))
    return dy.numpy()


def gather_rows(data, idx, ctx=None):
    ctx = ctx or get_context()
    data = np.ascontiguousarray(data, _f32)
    idx = np.ascontiguousarray(idx, np.int32)
    dd, di = _dev(ctx, data), DeviceArray.from_numpy(ctx, idx, np.int32)
    out = DeviceArray(ctx, (idx.size, data.shape[1]))
    check(lib.b200_gather_rows(ctx.h, C.c_int(idx.size), C.c_int(data.shape[1]), dd.ptr, di.ptr, out.ptr))
    return out.numpy()


# ---- SURVEY.md 8(f) kernels ------------------------------------------------------------------------
ACT.update({"log_logistic": 7, "softplus": 8, "softsign": 9, "leaky_relu": 10, "hardtanh": 11})


def _act_params(kind, leak=0.01, inf=-1.0, sup=1.0):
    return (leak, 0.0) if kind == "leaky_relu" else ((inf, sup) if kind == "hardtanh" else (0.0, 0.0))


def actf_fwd_ex(kind, x, ctx=None, **params):
    ctx = ctx or get_context()
    x = np.ascontiguousarray(x, _f32)
    p0, p1 = _act_params(kind, **params)
    dx, dy = _dev(ctx, x), DeviceArray(ctx, x.shape)
    check(lib.b200_actf_fwd_ex(ctx.h, C.c_int(ACT[kind]), C.c_float(p0), C.c_float(p1), C.c_size_t(x.size), dx.ptr, dy.ptr))
    return dy.numpy()


def actf_bwd_ex(kind, x, y, dy, ctx=None, **params):
    ctx = ctx or get_context()
    x, y, dy = (np.ascontiguousarray(a, _f32) for a in (x, y, dy))
    p0, p1 = _act_params(kind, **params)
    d_x, d_y, d_dy, d_dx = _dev(ctx, x), _dev(ctx, y), _dev(ctx, dy), DeviceArray(ctx, x.shape)
    check(lib.b200_actf_bwd_ex(ctx.h, C.c_int(ACT[kind]), C.c_float(p0), C.c_float(p1), C.c_size_t(x.size), d_x.ptr,
                               d_y.ptr, d_dy.ptr, d_dx.ptr))
    return d_dx.numpy()


def zero_one_loss(out, target, TH=0.5, ctx=None):
    ctx = ctx or get_context()
    out, target = np.ascontiguousarray(out, _f32), np.ascontiguousarray(target, _f32)
    M, Cc = out.shape
    d_o, d_t, rows = _dev(ctx, out), _dev(ctx, target), DeviceArray(ctx, (M,))
    check(lib.b200_zero_one_loss(ctx.h, C.c_int(M), C.c_int(Cc), d_o.ptr, d_t.ptr, C.c_int(target.shape[1]), C.c_float(TH),
                                 rows.ptr))
    return rows.numpy()


class DropoutStream:
    """The device-side MT19937 of the dropout component, seeded like `random(seed)`: mask(n, prob) returns
    the next n mask values of the reference's stream (dropout_component.cc:91-95)."""

    def __init__(self, seed, ctx=None):
        from . import random as _random
        self.ctx = ctx or get_context()
        lib.b200_mt_state_bytes.restype = C.c_size_t
        nbytes = lib.b200_mt_state_bytes()
        # same state the host generator has right after seeding: 624 words after the first reload, nothing read
        from ._lib import lib as _l
        r = _random(seed)
        words = (C.c_uint32 * 624)()
        nxt = C.c_int32()
        check(_l.b200h_random_export_state(r.h, words, C.byref(nxt)))
        host = np.zeros(nbytes // 4, dtype=np.uint32)
        host[:624] = np.frombuffer(words, dtype=np.uint32)
        host[624] = np.uint32(nxt.value)
        self.state = DeviceArray.from_numpy(self.ctx, host, np.uint32)

    def mask(self, n, prob):
        m = DeviceArray(self.ctx, (n,))
        check(lib.b200_dropout_mask(self.ctx.h, self.state.ptr, C.c_size_t(n), C.c_float(prob), m.ptr))
        return m.numpy()
