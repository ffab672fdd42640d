#!/usr/bin/env python
"""bench.py -- training throughput of the B200 hot path on BASELINE.json's headline config.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on host cores

Workload (BASELINE.json configs[1]): synthetic MLP 784-2048-2048-10, ReLU hidden layers,
log_softmax output + multi-class cross-entropy, bunch 1024 per GPU, SGD lr 0.01 / momentum 0.9 /
weight decay 1e-4 (0 on biases) / decay 1e-5, gradient smoothing on.  A "step" is one full
trainer:train_step (forward, loss, backward, weight gradients, scaling, SGD update).

One JSON line on stdout (rank 0).  `value` = samples/s summed over ranks, device-timed with CUDA
events per step (inputs resident in HBM, L2 flushed between steps outside the event bracket),
max over ranks.  `e2e` = the same through the host-pointer API with pinned host buffers: every
step copies its bunch H2D and its loss D2H inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TOPOLOGY = "784 inputs 2048 relu 2048 relu 10 log_softmax"
LAYERS = [(784, 2048), (2048, 2048), (2048, 10)]
BUNCH = 1024
METRIC = "MLP training samples/sec"
UNIT = "samples/s"


def step_flops(bunch):
    """Algorithmic FLOPs per step: forward 2*bs*P + weight gradient 2*bs*P +
    data gradient 2*bs*(P - P1): the network-input gradient is never needed for training."""
    p = sum(i * o for i, o in LAYERS)
    p1 = LAYERS[0][0] * LAYERS[0][1]
    return 2 * bunch * p + 2 * bunch * p + 2 * bunch * (p - p1)


def synthetic_bunch(seed, bunch):
    rng = np.random.RandomState(seed)
    x = rng.uniform(-1, 1, size=(bunch, LAYERS[0][0])).astype(np.float32)
    t = np.zeros((bunch, LAYERS[-1][1]), dtype=np.float32)
    t[np.arange(bunch), rng.randint(0, LAYERS[-1][1], size=bunch)] = 1.0
    return x, t


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_burst": p["bf16_tflops"], "bf16_sustained": p["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.samples = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.samples:
            f = [v.strip() for v in line.split(",")]
            if len(f) < 7:
                continue
            try:
                if t0 - 0.05 <= ts <= t1 + 0.15:
                    sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            if t0 - 0.05 <= ts <= t1 + 0.15:
                for nm, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------- reference arm
def oracle_trainer():
    from oracle import MTRand
    from oracle import april as A
    tr = A.SupervisedTrainer(A.mlp_all_all(TOPOLOGY), A.MultiClassCrossEntropy(), BUNCH).build()
    tr.set_option("learning_rate", 0.01)
    tr.set_option("momentum", 0.9)
    tr.set_option("weight_decay", 1e-4)
    tr.set_layerwise_option("b.", "weight_decay", 0)
    tr.randomize_weights(random=MTRand(1234), inf=-1, sup=1, use_fanin=True, use_fanout=True)
    return tr


def time_oracle(steps, warmup, budget_s=None):
    tr = oracle_trainer()
    x, t = synthetic_bunch(42, BUNCH)
    for _ in range(max(warmup, 1)):
        tr.train_step(x, t)
    t0 = time.time()
    done = 0
    for _ in range(steps):
        tr.train_step(x, t)
        done += 1
        if budget_s is not None and time.time() - t0 > budget_s:
            break
    dt = time.time() - t0
    return done * BUNCH / dt, dt / done, done


def cores_used():
    try:
        from threadpoolctl import threadpool_info
        n = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
        return int(n)
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, s_per_step, done = time_oracle(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "warmup": args.warmup, "ms_per_step": 1e3 * s_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "MLP %s bunch %d, SGD momentum .9 wd 1e-4" % (TOPOLOGY, BUNCH)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores_used(), "kind": "port",
                         "sample": "%d full train steps of the bunch-%d workload (numpy/OpenBLAS oracle)" % (done, BUNCH)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import ctypes as C
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="gloo")  # plumbing only: rendezvous, id exchange, max-over-ranks

    import april_ann_b200 as ann
    from april_ann_b200._lib import lib, check
    from april_ann_b200.parallel import init_data_parallel

    ctx = ann.get_context(local_rank)
    mode = ann.MATH_FP32 if args.math == "fp32" else ann.MATH_TF32
    ctx.set_math_mode(mode)
    tr = ann.trainable.supervised_trainer(ann.mlp.all_all.generate(TOPOLOGY), ann.loss.multi_class_cross_entropy(), BUNCH, ctx=ctx)
    tr.build()
    tr.set_option("learning_rate", 0.01)
    tr.set_option("momentum", 0.9)
    tr.set_option("weight_decay", 1e-4)
    tr.set_layerwise_option("b.", "weight_decay", 0)
    tr.randomize_weights(random=ann.random(1234), inf=-1, sup=1, use_fanin=True, use_fanout=True)
    if world > 1:
        init_data_parallel(tr, dist)

    # pinned host bunches (distinct per rank and rotated per step)
    nb = 4
    hx = C.c_void_p()
    ht = C.c_void_p()
    in_sz, out_sz = LAYERS[0][0], LAYERS[-1][1]
    check(lib.b200_host_alloc(C.byref(hx), C.c_size_t(nb * BUNCH * in_sz * 4)))
    check(lib.b200_host_alloc(C.byref(ht), C.c_size_t(nb * BUNCH * out_sz * 4)))
    px = np.ctypeslib.as_array(C.cast(hx, C.POINTER(C.c_float)), shape=(nb, BUNCH, in_sz))
    pt = np.ctypeslib.as_array(C.cast(ht, C.POINTER(C.c_float)), shape=(nb, BUNCH, out_sz))
    for i in range(nb):
        px[i], pt[i] = synthetic_bunch(42 + 100 * rank + i, BUNCH)
    hl = C.c_void_p()
    check(lib.b200_host_alloc(C.byref(hl), C.c_size_t(8 * 4096)))
    ploss = np.ctypeslib.as_array(C.cast(hl, C.POINTER(C.c_double)), shape=(4096,))

    # L2 flush buffer (2x the 126 MB L2)
    flush_bytes = 256 << 20
    fl = C.c_void_p()
    check(lib.b200_malloc(ctx.h, C.byref(fl), C.c_size_t(flush_bytes)))

    def flush():
        check(lib.b200_memset_zero(ctx.h, fl, C.c_size_t(flush_bytes)))

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()

    # ---- device-resident throughput -------------------------------------------------
    tr.stage(px[0], pt[0], BUNCH)
    for _ in range(max(args.warmup, 3)):
        tr.step_staged(BUNCH)
    ctx.sync()
    K = args.steps
    evs = []
    for _ in range(2 * K):
        e = C.c_void_p()
        check(lib.b200_event_create(C.byref(e)))
        evs.append(e)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    launches0 = ctx.launch_count()
    wall0 = time.time()
    for k in range(K):
        flush()
        check(lib.b200_event_record(ctx.h, evs[2 * k]))
        tr.step_staged(BUNCH)
        check(lib.b200_event_record(ctx.h, evs[2 * k + 1]))
    barrier()
    wall1 = time.time()
    launches = ctx.launch_count() - launches0
    dev_ms = 0.0
    per_step = []
    for k in range(K):
        ms = C.c_float()
        check(lib.b200_event_elapsed_ms(evs[2 * k], evs[2 * k + 1], C.byref(ms)))
        per_step.append(ms.value)
        dev_ms += ms.value
    # hot-L2 variant: K steps back to back inside one event bracket
    barrier()
    check(lib.b200_event_record(ctx.h, evs[0]))
    for k in range(K):
        tr.step_staged(BUNCH)
    check(lib.b200_event_record(ctx.h, evs[1]))
    barrier()
    ms = C.c_float()
    check(lib.b200_event_elapsed_ms(evs[0], evs[1], C.byref(ms)))
    hot_ms = ms.value
    clocks = sampler.stop(wall0, time.time()) if sampler else None

    # ---- end to end: pinned host -> device every step, loss back every step ------------
    # (own warm-up: the pipelined feed alternates two staging slots, each with its own captured step)
    for k in range(max(args.warmup, 6)):
        tr.stage(px[k % nb], pt[k % nb], BUNCH)
        tr.step_staged(BUNCH)
        check(lib.b200h_trainer_last_loss_async(tr.h, C.c_void_p(hl.value + 8 * (k % 4096))))
    barrier()
    t0 = time.time()
    check(lib.b200_event_record(ctx.h, evs[2]))
    for k in range(K):
        tr.stage(px[k % nb], pt[k % nb], BUNCH)
        tr.step_staged(BUNCH)
        check(lib.b200h_trainer_last_loss_async(tr.h, C.c_void_p(hl.value + 8 * (k % 4096))))
    check(lib.b200_event_record(ctx.h, evs[3]))
    barrier()
    e2e_wall = time.time() - t0
    check(lib.b200_event_elapsed_ms(evs[2], evs[3], C.byref(ms)))
    e2e_ms = ms.value
    last_loss = float(ploss[(K - 1) % 4096]) / BUNCH

    if dist is not None:
        import torch
        tt = torch.tensor([dev_ms, hot_ms, e2e_ms, e2e_wall * 1e3], dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms, hot_ms, e2e_ms, e2e_wall = tt[0].item(), tt[1].item(), tt[2].item(), tt[3].item() / 1e3

    if rank != 0:
        if dist is not None:
            dist.barrier()
        return

    pk = peaks()
    value = world * BUNCH * K / (dev_ms / 1e3)
    flops = step_flops(BUNCH)
    # tensor peak for the compute type: TF32 runs at half the bf16 rate on the tcgen05 pipe;
    # the denominator is the MEASURED bf16 cuBLAS number / 2.  fp32 (FFMA) mode has no tensor peak.
    tf32_peak = pk["bf16_sustained"] / 2.0
    step_tflops = flops / (dev_ms / K / 1e3) / 1e12

    roof = gemm_roofline(ann, ctx, lib, check, C, pk, args.math)
    cpu_val, cpu_s, cpu_done = time_oracle(40, 1, budget_s=15.0)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3),
        "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "tf32" if args.math == "tf32" else "f32", "data": "synthetic",
        "config": {"workload": "MLP %s, bunch %d per GPU, MCCE loss, SGD lr .01 momentum .9 wd 1e-4 (BASELINE configs[1])" % (TOPOLOGY, BUNCH),
                   "global_bunch": world * BUNCH, "parallelism": "dp%d" % world, "math": args.math,
                   "l2": "flushed between steps (256 MiB memset outside the per-step event bracket)",
                   "timing": "sum of per-step CUDA-event durations on the launching stream, max over ranks"},
        "clocks": clocks,
        "e2e": {"value": world * BUNCH * K / (e2e_ms / 1e3), "unit": UNIT,
                "h2d_bytes_per_step": BUNCH * (in_sz + out_sz) * 4, "d2h_bytes_per_step": 8,
                "wall_value": world * BUNCH * K / e2e_wall, "last_loss": last_loss},
        "gpu_launches": int(launches),
        "roofline": roof,
        "cpu_baseline": {"value": cpu_val, "unit": UNIT, "cores": cores_used(), "kind": "port",
                         "sample": "%d full train steps of the same bunch-%d workload on the host (numpy/OpenBLAS oracle)" % (cpu_done, BUNCH)},
        "step_tflops": step_tflops, "step_frac_of_tf32_peak": step_tflops / tf32_peak,
        "tf32_peak_tflops": tf32_peak, "peak_source": pk["source"] + "; tf32 = bf16_sustained/2",
        "hot_l2_value": world * BUNCH * K / (hot_ms / 1e3),
        "per_step_ms_minmax": [min(per_step), max(per_step)],
    }
    print(json.dumps(line))
    if dist is not None:
        dist.barrier()


def gemm_roofline(ann, ctx, lib, check, C, pk, math):
    """Dominant kernel: the contraction.  Times the largest GEMM of the step (forward of the
    2048x2048 layer: M=1024, N=2048, K=2048, bias + ReLU epilogue) alone: `reps` launches back to back
    inside one CUDA-event bracket on the launching stream, rotating over operand sets that together
    (8 x 32 MiB) exceed the 126 MB L2, so no launch finds its operands cached; achieved = 2*M*N*K / mean
    launch time."""
    from april_ann_b200.ops import DeviceArray
    M, N, K = BUNCH, 2048, 2048
    rng = np.random.RandomState(0)
    nsets = 8
    sets = []
    for i in range(nsets):
        X = DeviceArray.from_numpy(ctx, rng.uniform(-1, 1, (M, K)).astype(np.float32))
        W = DeviceArray.from_numpy(ctx, rng.uniform(-0.05, 0.05, (N, K)).astype(np.float32))
        sets.append((X, W, DeviceArray(ctx, (M, N))))
    b = DeviceArray.from_numpy(ctx, np.zeros(N, np.float32))
    e0, e1 = C.c_void_p(), C.c_void_p()
    check(lib.b200_event_create(C.byref(e0)))
    check(lib.b200_event_create(C.byref(e1)))

    def launch(i):
        X, W, Y = sets[i % nsets]
        check(lib.b200_linear_fwd(ctx.h, C.c_int(M), C.c_int(N), C.c_int(K), X.ptr, C.c_int(K), W.ptr, C.c_int(K),
                                  b.ptr, C.c_int(3), Y.ptr, C.c_int(N)))
    for i in range(2 * nsets):
        launch(i)
    ctx.sync()
    reps = 64
    check(lib.b200_event_record(ctx.h, e0))
    for i in range(reps):
        launch(i)
    check(lib.b200_event_record(ctx.h, e1))
    ms = C.c_float()
    check(lib.b200_event_elapsed_ms(e0, e1, C.byref(ms)))
    t = ms.value / reps / 1e3
    achieved = 2.0 * M * N * K / t / 1e12
    if math == "tf32":
        peak = pk["bf16_burst"] / 2.0
        note = "tf32 tensor peak = measured bf16 burst / 2 (%s)" % pk["source"]
    else:
        peak = 75.0
        note = "fp32 FFMA mode: nominal 75 TFLOP/s CUDA-core peak (no measured figure)"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    return {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "traffic": traffic, "kernel": "linear_fwd M=%d N=%d K=%d (+bias+relu epilogue), %s" % (M, N, K, math),
            "launch_us": t * 1e6, "note": note + "; %d launches back to back over %d operand sets (256 MiB > L2)" % (reps, nsets)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--math", default="tf32", choices=["tf32", "fp32"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
