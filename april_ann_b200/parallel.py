"""Data-parallel plumbing: one process per GPU, a weight replica per rank, gradients summed
with NCCL before the fused SGD update.  New relative to the reference (single device 0,
mathcore/c_src/gpu_helper.h:65-68).  torch.distributed is used only for the rendezvous, the
NCCL unique-id exchange and max-over-ranks timing; the collective on the data path is the
library's own ncclAllReduce on the flat gradient arena (csrc/nccl_dp.cu).

Semantics (kept equal to a single-device step on the concatenated bunch): every rank runs
forward/backward on its rows of the global bunch; raw gradients are sums over rows, so the sum
over ranks is the global sum; the reference's smoothing factor 1/sqrt(shared_count * bunch)
(packages/trainable/lua_src/supervised.lua:797-803) uses the GLOBAL bunch size.
"""
import ctypes as C
import math

from ._lib import lib, check


def shard_rows(n, rank, world):
    """Row range [lo, hi) of a global bunch of n rows owned by `rank` (contiguous, balanced)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_float4(n, rank, world):
    """Element range [lo, hi) of a tensor of n floats that `rank` owns in the peer-memory replica group:
    the tensor is cut in float4 units, ceil(ceil(n/4) / world) per rank (csrc/dp_fused.cu, `plan`); the
    last float4 may reach into the tensor's padding inside the arena, never past it."""
    n4 = (n + 3) // 4
    per = (n4 + world - 1) // world
    lo = min(n4, rank * per)
    hi = min(n4, lo + per)
    return min(n, 4 * lo), min(n, 4 * hi)


def dp_grad_scale(shared_count, global_bunch):
    """The reference's gradient smoothing with the global bunch size."""
    return 1.0 / math.sqrt(max(shared_count, 1) * global_bunch)


def exchange_unique_id(dist):
    """rank 0 creates the ncclUniqueId; everybody receives the 128 bytes."""
    buf = (C.c_char * 128)()
    if dist.get_rank() == 0:
        check(lib.b200_comm_unique_id(buf))
    obj = [bytes(buf)]
    dist.broadcast_object_list(obj, src=0)
    return obj[0]


def init_data_parallel(trainer, dist, fused=True, multicast=True):
    """Creates the NCCL communicator of the trainer's context, makes the trainer scale gradients
    by the global bunch and all-reduce them, and broadcasts rank 0's weights to every replica.
    fused: the update of a step runs as one kernel per bucket over NVLink peer memory -- through the NVSwitch's
    in-switch reduction (multicast / NVLS) when `multicast` and the fabric allow it, else with P2P loads and
    stores (CUDA IPC mappings)."""
    import os
    world, rank = dist.get_world_size(), dist.get_rank()
    uid = exchange_unique_id(dist)
    check(lib.b200_comm_init(trainer.ctx.h, C.c_int(world), C.c_int(rank), uid))
    trainer.set_data_parallel(world, rank)
    trainer.broadcast_weights()
    dist.barrier()
    if fused and 2 <= world <= 8:
        done = False
        # Transport of the fused update (measured on the pool's boxes, C2, profiles/dp_r2.md):
        #   2 GPUs: P2P loads / stores 175 us per step, in-switch reduction (NVLS multicast) 217 us
        #   8 GPUs: P2P 205 us, NVLS 201 us
        # so the multicast path is the default from 8 ranks on; B200_DP_SYMMETRIC=1 / 0 forces it on / off.
        want = os.environ.get("B200_DP_SYMMETRIC")
        use_symm = multicast and (world >= 8 if want is None else want != "0")
        if use_symm:
            done = connect_symmetric_memory(trainer, dist)
        if not done:
            connect_peer_memory(trainer, dist)
    dist.barrier()


def connect_symmetric_memory(trainer, dist):
    """Allocates this rank's [weights | gradients | flags] block in symmetric memory (torch's allocator: CUDA VMM
    allocations exchanged between the processes and bound to a multicast object -- plumbing only), hands the
    mapped addresses to the trainer.  Every rank must succeed, otherwise all of them use the IPC path."""
    world, rank = dist.get_world_size(), dist.get_rank()
    ok, keep, mc = True, None, 0
    # the rendezvous below is a collective: agree first that every rank can take part at all (imports, a device
    # that supports multicast objects), so that no rank waits inside it for one that never arrives
    try:
        import torch
        import torch.distributed._symmetric_memory as symm_mem
        from cuda import cuda as _cu
        _cu.cuInit(0)
        _, _dev = _cu.cuDeviceGet(trainer.ctx.device)
        _, _mcs = _cu.cuDeviceGetAttribute(_cu.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, _dev)
        pre = bool(_mcs) and torch.cuda.is_available()
    except Exception:
        pre = False
    agree = [None] * world
    dist.all_gather_object(agree, pre)
    if not all(agree):
        return False
    try:
        lib.b200h_trainer_dp_symmetric_bytes.restype = C.c_size_t
        nbytes = int(lib.b200h_trainer_dp_symmetric_bytes(trainer.h))
        if nbytes <= 0:
            raise RuntimeError("trainer not built")
        dev = torch.device("cuda", trainer.ctx.device)
        torch.cuda.set_device(dev)
        buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=dev)
        hdl = symm_mem.rendezvous(buf, dist.group.WORLD.group_name)
        bases = [int(p) for p in hdl.buffer_ptrs]
        mc = int(hdl.multicast_ptr) if getattr(hdl, "multicast_ptr", 0) else 0
        keep = (buf, hdl)
        if len(bases) != world or not all(bases):
            raise RuntimeError("incomplete peer mapping")
    except Exception:
        ok = False
    flags = [None] * world
    dist.all_gather_object(flags, (ok, bool(ok and mc)))
    if not all(f[0] for f in flags):
        return False
    if not all(f[1] for f in flags):
        mc = 0     # some rank has no multicast mapping: everybody uses P2P loads / stores on the symmetric buffers
    torch.cuda.synchronize()
    arr = (C.c_void_p * world)(*bases)
    try:
        check(lib.b200h_trainer_dp_connect_symmetric(trainer.h, C.c_int(world), C.c_int(rank), arr, C.c_void_p(mc or None),
                                                     C.c_size_t(nbytes)))
    except Exception:
        ok = False
    dist.all_gather_object(flags, ok)
    if not all(flags):
        if ok:
            trainer.set_flag("dp_fused", 0)
        return False
    trainer._symmetric = keep       # the allocation lives as long as the trainer
    trainer.dp_transport = "multicast" if mc else "symmetric-p2p"
    return True


def connect_peer_memory(trainer, dist):
    """Maps every rank's weight arena, gradient arena and flag block into this process (CUDA IPC; the
    64-byte handles travel over the host-side rendezvous) so that the update of a step runs as one
    reduce-scatter + SGD + all-gather kernel over NVLink (csrc/dp_fused.cu).  Every rank must succeed,
    otherwise all of them stay on the NCCL all-reduce path."""
    world, rank = dist.get_world_size(), dist.get_rank()
    mine = (C.c_ubyte * 256)()
    ok = True
    try:
        check(lib.b200h_trainer_dp_export(trainer.h, C.c_int(world), mine))
    except Exception:
        ok = False
    gathered = [None] * world
    dist.all_gather_object(gathered, (ok, bytes(mine)))
    if not all(g[0] for g in gathered):
        return False
    blob = b"".join(g[1] for g in gathered)
    buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
    try:
        check(lib.b200h_trainer_dp_connect(trainer.h, C.c_int(world), C.c_int(rank), buf))
    except Exception:
        ok = False
    flags = [None] * world
    dist.all_gather_object(flags, ok)
    if not all(flags):
        # some rank could not map a peer (no P2P between the devices, IPC not permitted ...): every rank goes
        # back to the NCCL all-reduce path, the replicas must not disagree about the protocol
        if ok:
            trainer.set_flag("dp_fused", 0)
        return False
    return True
