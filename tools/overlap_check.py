"""Does the NCCL all-reduce on the communication stream overlap with tcgen05 GEMMs on the compute
stream?  (torchrun, one rank per GPU)  Times 8 GEMMs alone, 4 all-reduces alone, and both forked."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch.distributed as dist  # noqa: E402

import april_ann_b200 as ann  # noqa: E402
from april_ann_b200._lib import lib, check  # noqa: E402
from april_ann_b200.ops import DeviceArray  # noqa: E402
from april_ann_b200.parallel import exchange_unique_id  # noqa: E402

os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group(backend="gloo")
rank, world = dist.get_rank(), dist.get_world_size()
ctx = ann.get_context(int(os.environ.get("LOCAL_RANK", "0")))
ctx.set_math_mode(ann.MATH_TF32)
uid = exchange_unique_id(dist)
check(lib.b200_comm_init(ctx.h, C.c_int(world), C.c_int(rank), uid))
e0, e1 = C.c_void_p(), C.c_void_p()
check(lib.b200_event_create(C.byref(e0)))
check(lib.b200_event_create(C.byref(e1)))
I = C.c_int
M, N, K = 1024, 2048, 2048
rng = np.random.RandomState(0)
A = DeviceArray.from_numpy(ctx, rng.uniform(-1, 1, (M, K)).astype(np.float32))
B = DeviceArray.from_numpy(ctx, rng.uniform(-1, 1, (N, K)).astype(np.float32))
Cm = DeviceArray(ctx, (M, N))
n = int(17e6 / 4)
buf = DeviceArray(ctx, (n,))
buf.zero()


def gemms():
    for _ in range(8):
        check(lib.b200_sgemm(ctx.h, I(0), I(1), I(M), I(N), I(K), C.c_float(1.0), A.ptr, I(K), B.ptr, I(K), C.c_float(0.0), Cm.ptr, I(N)))


def ars_sync():
    for _ in range(4):
        check(lib.b200_allreduce_sum(ctx.h, buf.ptr, C.c_size_t(n)))


def ars_async():
    for s in range(4):
        check(lib.b200_allreduce_sum_async(ctx.h, buf.ptr, C.c_size_t(n), I(s)))


def timed(fn):
    ctx.sync()
    dist.barrier()
    check(lib.b200_event_record(ctx.h, e0))
    fn()
    check(lib.b200_event_record(ctx.h, e1))
    ms = C.c_float()
    check(lib.b200_event_elapsed_ms(e0, e1, C.byref(ms)))
    return ms.value * 1e3


def both():
    ars_async()          # forked onto the comm stream
    gemms()              # compute stream
    for s in range(4):
        check(lib.b200_comm_wait(ctx.h, I(s)))


for _ in range(3):
    gemms(); ars_sync(); both()
tg = min(timed(gemms) for _ in range(5))
ta = min(timed(ars_sync) for _ in range(5))
tb = min(timed(both) for _ in range(5))
if rank == 0:
    print("8 GEMMs alone %.1f us | 4 all-reduces (17 MB) alone %.1f us | forked together %.1f us (sum %.1f, max %.1f)" % (
        tg, ta, tb, tg + ta, max(tg, ta)))
dist.barrier()
dist.destroy_process_group()
