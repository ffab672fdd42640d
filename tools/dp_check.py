"""Data-parallel parity on real GPUs (launch with torchrun, one rank per GPU):
k train steps with the global bunch sharded over the ranks (NCCL bucketed all-reduce overlapped
with the backward pass) must give the weights of k single-GPU steps on the whole bunch, within
fp32 summation-order noise.  Rank 0 also runs the single-GPU reference.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_check.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch.distributed as dist  # noqa: E402

import april_ann_b200 as ann  # noqa: E402
from april_ann_b200.parallel import init_data_parallel, shard_rows  # noqa: E402

TOPO = "784 inputs 2048 relu 2048 relu 10 log_softmax"
GB = 1024


def make(ctx):
    tr = ann.trainable.supervised_trainer(ann.mlp.all_all.generate(TOPO), ann.loss.multi_class_cross_entropy(), GB, ctx=ctx)
    tr.build()
    tr.set_option("learning_rate", 0.01)
    tr.set_option("momentum", 0.9)
    tr.set_option("weight_decay", 1e-4)
    tr.set_layerwise_option("b.", "weight_decay", 0)
    tr.randomize_weights(random=ann.random(1234), inf=-1, sup=1, use_fanin=True, use_fanout=True)
    return tr


def main():
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group(backend="gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    ctx = ann.get_context(int(os.environ.get("LOCAL_RANK", "0")))
    ctx.set_math_mode(ann.MATH_FP32)
    rng = np.random.RandomState(5)
    x = rng.uniform(-1, 1, (GB, 784)).astype(np.float32)
    t = np.zeros((GB, 10), np.float32)
    t[np.arange(GB), rng.randint(0, 10, GB)] = 1.0
    tr = make(ctx)
    init_data_parallel(tr, dist)
    lo, hi = shard_rows(GB, rank, world)
    losses = []
    for _ in range(6):
        l, _ = tr.train_step(x[lo:hi], t[lo:hi])
        losses.append(l)
    ws = {n: tr.weights(n) for n in tr.weight_names()}
    ok = True
    if rank == 0:
        ref = make(ctx)
        ref_losses = [ref.train_step(x, t)[0] for _ in range(6)]
        for n in ref.weight_names():
            a, b = ws[n].astype(np.float64), ref.weights(n).astype(np.float64)
            err = np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)
            print("%-4s rel-L2 vs single GPU on the global bunch: %.3e" % (n, err))
            ok = ok and err < 3e-5   # six momentum steps of fp32 summation-order differences (sum over ranks vs sum over chunks)
        print("single-GPU losses", ["%.5f" % v for v in ref_losses])
    print("rank %d shard losses %s" % (rank, ["%.5f" % v for v in losses]))
    dist.barrier()
    if rank == 0:
        print("DP PARITY", "OK" if ok else "FAILED")
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
