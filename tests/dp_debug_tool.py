"""Per-tensor parity of the data-parallel update against the oracle, step by step (GPU box, torchrun).
    python -m torch.distributed.run --nproc-per-node N tests/dp_debug_tool.py [fp32|tf32] [graph|nograph] [steps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch.distributed as dist  # noqa: E402

os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group(backend="gloo")
rank, world = dist.get_rank(), dist.get_world_size()
import april_ann_b200 as ann  # noqa: E402
from april_ann_b200 import configs as CFG  # noqa: E402
from april_ann_b200.parallel import init_data_parallel  # noqa: E402
from oracle import configs as OC  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
graph = (sys.argv[2] if len(sys.argv) > 2 else "graph") == "graph"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
name = os.environ.get("CFGNAME", "C2")
ctx = ann.get_context(int(os.environ.get("LOCAL_RANK", "0")))
ctx.set_math_mode(ann.MATH_TF32 if mode == "tf32" else ann.MATH_FP32)
tr = CFG.build_trainer(ann, name, ctx=ctx)
tr.set_flag("cuda_graph", int(graph))
tr.randomize_weights(random=ann.random(1234), inf=-1, sup=1, use_fanin=True, use_fanout=True)
init_data_parallel(tr, dist)
pb = CFG.CONFIGS[name]["bunch"]
names = tr.weight_names()
w0 = {n: tr.weights(n) for n in names}
ref = None
if rank == 0:
    ref = OC.build_trainer(name, world * pb)
    for n in names:
        ref.weights[n][...] = w0[n].reshape(ref.weights[n].shape)
for s in range(steps):
    x, t = CFG.synthetic_bunch(name, 9000 + 1000 * s + rank, pb)
    l, _ = tr.train_step(x, t, bunch_size=pb)
    ctx.sync()
    dist.barrier()
    w1 = {n: tr.weights(n) for n in names}
    if rank == 0:
        xs, ts = zip(*[CFG.synthetic_bunch(name, 9000 + 1000 * s + r, pb) for r in range(world)])
        before = {n: ref.weights[n].copy() for n in names}
        lr_, _ = ref.train_step(np.concatenate(xs), np.concatenate(ts), bunch_size=world * pb)
        msg = []
        for n in names:
            dr = ref.weights[n].astype(np.float64) - before[n]
            dg = w1[n].astype(np.float64).reshape(dr.shape) - before[n]
            msg.append("%s %.1e" % (n, np.linalg.norm(dg - dr) / max(np.linalg.norm(dr), 1e-30)))
            ref.weights[n][...] = w1[n].reshape(ref.weights[n].shape)     # continue from the device's weights
        print("step %d loss gpu(rank0) %.6f oracle %.6f | step-movement rel-L2: %s" % (s, l, lr_, "  ".join(msg)), flush=True)
    dist.barrier()
