"""NCCL all-reduce latency of the gradient buckets (torchrun, one rank per GPU)."""
import ctypes as C
import os
import sys

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
uid = exchange_unique_id(dist)
check(lib.b200_comm_init(ctx.h, C.c_int(world), C.c_int(rank), uid))
e0, e1 = C.c_void_p(), C.c_void_p()
check(lib.b200_event_create(C.byref(e0)))
check(lib.b200_event_create(C.byref(e1)))
for mb in (0.08, 1, 6.4, 17, 23, 64, 537):
    n = int(mb * 1e6 / 4)
    buf = DeviceArray(ctx, (n,))
    buf.zero()
    for _ in range(5):
        check(lib.b200_allreduce_sum(ctx.h, buf.ptr, C.c_size_t(n)))
    ctx.sync()
    dist.barrier()
    check(lib.b200_event_record(ctx.h, e0))
    for _ in range(20):
        check(lib.b200_allreduce_sum(ctx.h, buf.ptr, C.c_size_t(n)))
    check(lib.b200_event_record(ctx.h, e1))
    ms = C.c_float()
    check(lib.b200_event_elapsed_ms(e0, e1, C.byref(ms)))
    us = ms.value / 20 * 1e3
    if rank == 0:
        print("all-reduce %7.2f MB x%d ranks: %8.1f us  algbw %6.1f GB/s  busbw %6.1f GB/s" % (
            mb, world, us, mb * 1e6 / us / 1e3, mb * 1e6 / us / 1e3 * 2 * (world - 1) / world), flush=True)
    buf.free()
dist.barrier()
dist.destroy_process_group()
