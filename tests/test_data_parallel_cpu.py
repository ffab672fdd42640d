"""Host-side data-parallel logic on CPU: world_size-2 gloo group (no GPU).  Each rank runs the
oracle's forward/backward on its rows of a global bunch, raw gradient sums are all-reduced, and
the reference's smoothing factor is applied with the GLOBAL bunch size
(april_ann_b200/parallel.py) -- the result must equal one single-process step on the whole bunch,
which is the property the NCCL path relies on."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

TOPO = "12 inputs 9 tanh 7 relu 4 log_softmax"
GLOBAL_BUNCH = 10


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_trainer():
    from oracle import MTRand
    from oracle import april as A
    tr = A.SupervisedTrainer(A.mlp_all_all(TOPO), A.MultiClassCrossEntropy(), GLOBAL_BUNCH).build()
    tr.set_option("learning_rate", 0.1)
    tr.set_option("momentum", 0.5)
    tr.set_option("weight_decay", 1e-3)
    tr.set_layerwise_option("b.", "weight_decay", 0)
    tr.randomize_weights(random=MTRand(1234), inf=-1, sup=1, use_fanin=True, use_fanout=True)
    return tr


def _data():
    from oracle import MTRand
    r = MTRand(99)
    x = (r.rand_array(GLOBAL_BUNCH * 12, 2.0) - 1.0).astype(np.float32).reshape(GLOBAL_BUNCH, 12)
    t = np.zeros((GLOBAL_BUNCH, 4), dtype=np.float32)
    t[np.arange(GLOBAL_BUNCH), [r.randInt(0, 3) for _ in range(GLOBAL_BUNCH)]] = 1.0
    return x, t


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    # the package needs the built library to import; only its host-side helpers are used here
    from april_ann_b200.parallel import shard_rows, dp_grad_scale
    tr = _make_trainer()
    x, t = _data()
    lo, hi = shard_rows(GLOBAL_BUNCH, rank, world)
    for _ in range(3):
        out = tr.net.forward(x[lo:hi], True)
        _, rows = tr.loss.compute_loss(out, t[lo:hi])
        tr.net.backprop(tr.loss.gradient(out, t[lo:hi]))
        grads, counts = {}, {}
        tr.net.compute_gradients(grads, counts)
        for name in sorted(grads):
            g = torch.from_numpy(grads[name])
            dist.all_reduce(g, op=dist.ReduceOp.SUM)          # raw sums over rows -> global sum
            grads[name] *= np.float32(dp_grad_scale(counts.get(name, 0) or 1, GLOBAL_BUNCH))
        tr.optimizer.execute(tr.weights, grads)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), **tr.weights)
    dist.barrier()
    dist.destroy_process_group()


def _worker_sharded(rank, world, port, out_dir):
    """The protocol of csrc/dp_fused.cu with gloo standing in for NVLink: every rank reduces only ITS
    shard of every tensor (fixed rank order), applies the SGD step to that shard only (it alone keeps the
    shard's momentum), and the updated shards are gathered into every replica."""
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from april_ann_b200.parallel import shard_rows, dp_grad_scale, shard_float4
    tr = _make_trainer()
    x, t = _data()
    lo, hi = shard_rows(GLOBAL_BUNCH, rank, world)
    for _ in range(3):
        out = tr.net.forward(x[lo:hi], True)
        tr.net.backprop(tr.loss.gradient(out, t[lo:hi]))
        grads, counts = {}, {}
        tr.net.compute_gradients(grads, counts)
        new_w = {}
        for name in sorted(grads):
            flat = np.ascontiguousarray(grads[name]).reshape(-1)
            peers = [torch.empty(flat.size, dtype=torch.float32) for _ in range(world)]
            dist.all_gather(peers, torch.from_numpy(flat.copy()))        # "peer memory": everybody's gradients
            a, b = shard_float4(flat.size, rank, world)
            g = np.zeros(b - a, dtype=np.float32)
            for p in range(world):                                        # fixed rank order
                g += peers[p].numpy()[a:b]
            g *= np.float32(dp_grad_scale(counts.get(name, 0) or 1, GLOBAL_BUNCH))
            # SGD on the shard only: views into the trainer's weights / momentum
            w_sh = tr.weights[name].reshape(-1)[a:b]
            sub_w, sub_g = {name: w_sh}, {name: g}
            if name not in tr.optimizer.update:
                tr.optimizer.update[name] = np.zeros_like(tr.weights[name])
            full_u = tr.optimizer.update[name]
            tr.optimizer.update[name] = full_u.reshape(-1)[a:b]
            count_before = tr.optimizer.count
            tr.optimizer.execute(sub_w, sub_g)
            tr.optimizer.count = count_before                             # one step counter for all tensors
            full_u.reshape(-1)[a:b] = tr.optimizer.update[name]
            tr.optimizer.update[name] = full_u
            new_w[name] = (a, b, w_sh.copy())
        tr.optimizer.count += 1
        for name in sorted(new_w):                                        # all-gather of the updated shards
            a, b, w_sh = new_w[name]
            spans = [None] * world
            dist.all_gather_object(spans, (a, b, w_sh))
            flat_w = tr.weights[name].reshape(-1)
            for (pa, pb, pw) in spans:
                flat_w[pa:pb] = pw
    np.savez(os.path.join(out_dir, "shard_rank%d.npz" % rank), **tr.weights)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_float4_covers_every_tensor_once():
    from april_ann_b200.parallel import shard_float4
    for n in (1, 3, 4, 10, 127, 128, 2048, 20480, 1605632):
        for world in (2, 3, 4, 8):
            spans = [shard_float4(n, r, world) for r in range(world)]
            covered = np.zeros(n, dtype=np.int32)
            for a, b in spans:
                assert 0 <= a <= b <= n and (a % 4 == 0 or a == b)   # shards start on a float4 (empty ones aside)
                covered[a:b] += 1
            assert (covered == 1).all(), (n, world, spans)


def test_two_rank_sharded_update_equals_single_process_step(tmp_path):
    """reduce-scatter + SGD on the owner's shard + all-gather == all-reduce + SGD everywhere == one step on
    the global bunch (the property the peer-memory replica group relies on)."""
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker_sharded, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ref = _make_trainer()
    x, t = _data()
    for _ in range(3):
        ref.train_step(x, t)
    w0 = np.load(os.path.join(str(tmp_path), "shard_rank0.npz"))
    w1 = np.load(os.path.join(str(tmp_path), "shard_rank1.npz"))
    for name in ref.weights:
        assert np.array_equal(w0[name], w1[name]), name
        err = np.abs(w0[name] - ref.weights[name]).max()
        assert err < 2e-6, (name, err)


def test_shard_rows_partitions_every_bunch():
    from april_ann_b200.parallel import shard_rows
    for n in (1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 8):
            spans = [shard_rows(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_step_equals_single_process_step(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ref = _make_trainer()
    x, t = _data()
    for _ in range(3):
        ref.train_step(x, t)
    w0 = np.load(os.path.join(str(tmp_path), "rank0.npz"))
    w1 = np.load(os.path.join(str(tmp_path), "rank1.npz"))
    for name in ref.weights:
        assert np.array_equal(w0[name], w1[name]), name            # replicas stay identical
        err = np.abs(w0[name] - ref.weights[name]).max()
        assert err < 2e-6, (name, err)                              # == one step on the global bunch


# ---- bucket plan of the fused update (pure host logic behind b200h_dp_bucket_plan) ----------------------------

def _bucket_plan(sizes, bucket_bytes):
    import ctypes as C
    lib = C.CDLL(os.path.join(ROOT, "april_ann_b200", "libb200ann.so"))
    lib.b200h_dp_bucket_plan.argtypes = [C.POINTER(C.c_size_t), C.c_int, C.c_size_t, C.POINTER(C.c_int),
                                         C.POINTER(C.c_int), C.c_int]
    a = (C.c_size_t * max(len(sizes), 1))(*sizes)
    lo, hi = (C.c_int * 64)(), (C.c_int * 64)()
    n = lib.b200h_dp_bucket_plan(a, len(sizes), bucket_bytes, lo, hi, 64)
    return n, [(lo[i], hi[i]) for i in range(max(n, 0))]


def _check_plan(sizes, plan):
    # contiguous cover of [0, n), at most 32 tensors per bucket (the kernel's shared tables), at most 16 buckets
    assert plan[0][0] == 0 and plan[-1][1] == len(sizes)
    assert all(a[1] == b[0] for a, b in zip(plan, plan[1:]))
    assert all(0 < hi - lo <= 32 for lo, hi in plan)
    assert len(plan) <= 16


def test_bucket_plan_c2_two_large_tensors():
    mb = 1 << 20
    sizes = [8192, 2048 * 10 * 4, 8192, 2048 * 2048 * 4, 8192, 784 * 2048 * 4]
    n, plan = _bucket_plan(sizes, 4 * mb)
    _check_plan(sizes, plan)
    assert n == len(plan) == 2          # each closes on its large matrix
    assert all(sum(sizes[lo:hi]) >= 4 * mb for lo, hi in plan)


def test_bucket_plan_deep_net_of_small_layers_is_cut_at_32_tensors():
    # 20 layers of 128x128 (+ biases): 40 tensors, 1.3 MB -- one 4 MB bucket would exceed the kernel's 32-tensor
    # tables (this used to fail every step with B200_ERR_BAD_ARG once peer memory was connected)
    sizes = [128 * 128 * 4, 512] * 20
    n, plan = _bucket_plan(sizes, 4 << 20)
    _check_plan(sizes, plan)
    assert n == 2 and plan[0] == (0, 32)


def test_bucket_plan_threshold_doubles_until_sixteen_buckets_suffice():
    sizes = [1 << 20] * 100            # 100 tensors of 1 MB with a 1 MB threshold would be 100 buckets
    n, plan = _bucket_plan(sizes, 1 << 20)
    _check_plan(sizes, plan)
    assert 7 <= n <= 16


def test_bucket_plan_gives_up_beyond_512_tensors():
    assert _bucket_plan([64] * 600, 1)[0] == -1          # the trainer then takes the NCCL all-reduce path
    n, plan = _bucket_plan([64] * 512, 1)
    _check_plan([64] * 512, plan)
    assert n == 16
