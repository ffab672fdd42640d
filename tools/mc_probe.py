"""Probe: is NVLS multicast memory available to a torchrun job on this box (torch symmetric memory)?"""
import os
import sys

import torch
import torch.distributed as dist

os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
backend = sys.argv[1] if len(sys.argv) > 1 else "nccl"
dist.init_process_group(backend=backend, device_id=torch.device("cuda", local) if backend == "nccl" else None)
from cuda import cuda as cu  # noqa: E402
(err,) = cu.cuInit(0)
err, dev = cu.cuDeviceGet(local)
err, mc = cu.cuDeviceGetAttribute(cu.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev)
err, fab = cu.cuDeviceGetAttribute(cu.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED, dev)
print("rank", rank, "multicast supported", mc, "fabric handles", fab, flush=True)
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty(1 << 20, dtype=torch.float32, device=torch.device("cuda", local))
    hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
    print("rank", rank, "symm ok: multicast_ptr", hex(hdl.multicast_ptr), "buffers", [hex(p) for p in hdl.buffer_ptrs][:3],
          "signal pads", [hex(p) for p in hdl.signal_pad_ptrs][:2], "world", hdl.world_size, flush=True)
    # functional check of the switch reduction through torch's own op
    t.fill_(rank + 1.0)
    dist.barrier()
    torch.cuda.synchronize()
    try:
        torch.ops.symm_mem.multimem_all_reduce_(t, "sum", dist.group.WORLD.group_name)
        torch.cuda.synchronize()
        print("rank", rank, "multimem_all_reduce_ ->", float(t[0]), flush=True)
    except Exception as e:  # noqa: BLE001
        print("rank", rank, "multimem op failed:", repr(e)[:200], flush=True)
except Exception as e:  # noqa: BLE001
    print("rank", rank, "symmetric memory failed:", repr(e)[:400], flush=True)
dist.barrier()
dist.destroy_process_group()
