"""The fused data-parallel update (gradient reduce-scatter + SGD + weight all-gather over NVLink peer
memory, csrc/dp_fused.cu) alone, over all 5.8 M parameters of the C2 network (torchrun, one rank per GPU):
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/dp_bench.py"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch.distributed as dist  # noqa: E402

import april_ann_b200 as ann  # noqa: E402
from april_ann_b200._lib import lib, check  # noqa: E402
from april_ann_b200.parallel import init_data_parallel  # noqa: E402

os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group(backend="gloo")
rank, world = dist.get_rank(), dist.get_world_size()
ctx = ann.get_context(int(os.environ.get("LOCAL_RANK", "0")))
topo = os.environ.get("TOPOLOGY", "784 inputs 2048 relu 2048 relu 10 log_softmax")
tr = ann.trainable.supervised_trainer(ann.mlp.all_all.generate(topo), ann.loss.multi_class_cross_entropy(), 1024, ctx=ctx)
tr.build()
tr.set_option("learning_rate", 0.01)
tr.set_option("momentum", 0.9)
tr.set_option("weight_decay", 1e-4)
tr.randomize_weights(random=ann.random(1234), inf=-1, sup=1, use_fanin=True, use_fanout=True)
init_data_parallel(tr, dist)
us = C.c_float()
dist.barrier()
check(lib.b200h_trainer_dp_bench(tr.h, C.c_int(30), C.byref(us)))
n = tr.num_parameters() if hasattr(tr, "num_parameters") else 5824522
mb = 4.0 * n / 1e6
print("rank %d: fused update of %.1f MB of parameters over %d ranks: %.1f us  (gradient bytes / time = %.0f GB/s)" % (
    rank, mb, world, us.value, mb * 1e6 / us.value / 1e3), flush=True)
st = (C.c_longlong * 64)()
check(lib.b200h_trainer_dp_debug(tr.h, st))
print("rank %d bucket 0 (ns): wait for peers %d, shard update %d, publish %d" % (rank, st[1] - st[0], st[2] - st[1], st[3] - st[2]), flush=True)
dist.barrier()
dist.destroy_process_group()
