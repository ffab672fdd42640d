"""ctypes loader for libb200ann.so (the C ABI of include/b200ann.h + b200ann_host.h).

The product has no CPU path: if the CUDA library is missing this module raises at
import time, and creating a context without a B200 raises B200Error.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200ann.so")


class B200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("[b200 status %d] %s" % (code, msg))
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "april_ann_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C april_ann_b200/csrc). There is no CPU fallback." % LIB_PATH)
    # NCCL ships inside the torch wheel; make it findable for the lazy dlopen in nccl_dp.cu
    if "B200_NCCL_LIB" not in os.environ:
        try:
            import nvidia.nccl  # noqa: F401
            cand = os.path.join(os.path.dirname(nvidia.nccl.__file__), "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["B200_NCCL_LIB"] = cand
        except Exception:
            pass
    return C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)


lib = _load()

c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int)
c_void_pp = C.POINTER(C.c_void_p)

lib.b200_last_error_string.restype = C.c_char_p
lib.b200_stream.restype = C.c_void_p
lib.b200_stream.argtypes = [C.c_void_p]
for _n in ("b200h_random_new", "b200h_stack_new", "b200h_hyperplane_new", "b200h_dot_product_new",
           "b200h_bias_new", "b200h_actf_new", "b200h_actf_new_ex", "b200h_prelu_new", "b200h_dropout_new", "b200h_rewrap_new", "b200h_flatten_new",
           "b200h_convolution_new", "b200h_convolution_bias_new", "b200h_max_pooling_new",
           "b200h_mlp_generate", "b200h_trainer_new"):
    getattr(lib, _n).restype = C.c_void_p
lib.b200h_random_rand.restype = C.c_double
lib.b200h_random_rand.argtypes = [C.c_void_p, C.c_double]
lib.b200h_random_randint.restype = C.c_uint32
lib.b200h_random_randint.argtypes = [C.c_void_p, C.c_uint32]
lib.b200h_random_new.argtypes = [C.c_uint32]
lib.b200h_random_free.argtypes = [C.c_void_p]
lib.b200h_component_free.argtypes = [C.c_void_p]
lib.b200h_trainer_free.argtypes = [C.c_void_p]
lib.b200h_random_free.restype = None
lib.b200h_component_free.restype = None
lib.b200h_trainer_free.restype = None


def check(status):
    if status != 0:
        raise B200Error(status, lib.b200_last_error_string().decode("utf-8", "replace"))


def nonnull(ptr):
    if not ptr:
        raise B200Error(128, lib.b200_last_error_string().decode("utf-8", "replace"))
    return C.c_void_p(ptr)  # keep 64-bit handles intact when passed back without argtypes


def fptr(a):
    return a.ctypes.data_as(c_float_p)


def b(s):
    return s.encode() if s is not None else None
