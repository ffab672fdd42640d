#!/usr/bin/env python
"""bench.py -- training throughput of the B200 hot path on BASELINE.json's workloads.

    python bench.py --gpus N --steps K --warmup W [--config C2]      # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W            # the reference algorithm on host cores

Default workload = BASELINE.json configs[1] (C2): synthetic MLP 784-2048-2048-10, ReLU hidden layers,
log_softmax output + multi-class cross-entropy, bunch 1024 per GPU, SGD lr 0.01 / momentum 0.9 / weight decay
1e-4 (0 on biases) / decay 1e-5, gradient smoothing on.  `--config C1|C3|C4|C5` selects the other BASELINE
workloads (april_ann_b200/configs.py).  A "step" is one full trainer:train_step (forward, loss, backward,
weight gradients, scaling, update).

One JSON line on stdout (rank 0).  `value` = samples/s summed over ranks, device-timed with CUDA events per
step (inputs resident in HBM, L2 flushed between steps outside the event bracket), max over ranks.  `e2e` =
the same through the host-pointer API with pinned host buffers: every step copies its bunch H2D and its loss
D2H inside the timed region.  `parity` = after the timed loops every rank restarts from the same seeded
weights, runs k steps on fixed bunches, and rank 0 compares the result with the CPU oracle's k steps on the
concatenated global bunch (the oracle is the checker here, never the thing measured).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time


def physical_cores():
    try:
        avail = len(os.sched_getaffinity(0))
    except Exception:
        avail = os.cpu_count() or 1
    try:
        import psutil
        phys = psutil.cpu_count(logical=False) or avail
    except Exception:
        phys = avail
    return max(1, min(avail, phys))


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--math", default="tf32", choices=["tf32", "fp32"])
    ap.add_argument("--config", default="C2", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle parity phase")
    return ap.parse_args()


ARGS = parse_args()
# The CPU legs (reference arm, cpu_baseline, parity checker) use every physical core whatever the launcher
# exported: torch.distributed.run sets OMP_NUM_THREADS=1 for its children, which throttled the round-1
# reference arm to one thread at N >= 2.  Must happen before numpy loads its BLAS.
NCORES = physical_cores()
for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ[_v] = str(NCORES)

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import importlib.util  # noqa: E402

# shapes and builders of the workloads: loaded by path so that the reference arm never loads the CUDA library
_spec = importlib.util.spec_from_file_location("b200_configs", os.path.join(ROOT, "april_ann_b200", "configs.py"))
CFG = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(CFG)

METRIC = "MLP training samples/sec"
UNIT = "samples/s"
# bounded CPU samples: rows per oracle step (the full bunch where a step takes well under a second)
CPU_SAMPLE_BUNCH = {"C1": 32, "C2": 1024, "C3": 256, "C4": 128, "C5": 1024}
PARITY_BUNCH = {"C1": 32, "C2": 1024, "C3": 256, "C4": 128, "C5": 1024}
PARITY_STEPS = {"C1": 3, "C2": 3, "C3": 1, "C4": 3, "C5": 3}


def blas_threads(n):
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=n)
    except Exception:
        import contextlib
        return contextlib.nullcontext()


def cores_used():
    try:
        from threadpoolctl import threadpool_info
        n = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
        return int(n)
    except Exception:
        return NCORES


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


# --------------------------------------------------------------------------- CPU legs (oracle)
def oracle_trainer(name, bunch, weights=None):
    """The numpy port of the reference's CPU path on workload `name`; weights seeded like the GPU arm
    (random(1234), sorted-name order) unless given."""
    from oracle import MTRand
    from oracle import configs as OC
    tr = OC.build_trainer(name, bunch)
    if weights is None:
        if name == "C3":     # 134 M draws through the Python MT19937 would take minutes: any fixed values do for timing
            rng = np.random.RandomState(1234)
            for n in tr.weights_order:
                w = tr.weights[n]
                w[...] = rng.uniform(-1, 1, size=w.shape).astype(np.float32) / np.sqrt(sum(w.shape))
        else:
            tr.randomize_weights(random=MTRand(1234), inf=-1, sup=1, use_fanin=True, use_fanout=True)
    else:
        for n in tr.weights_order:
            tr.weights[n][...] = weights[n].reshape(tr.weights[n].shape)
    return tr


def time_oracle(name, steps, warmup, budget_s):
    """-> (samples/s, s per step, steps done, rows per step).  Bounded: rows per step = CPU_SAMPLE_BUNCH and the
    loop stops at `budget_s` seconds."""
    rows = CPU_SAMPLE_BUNCH[name]
    with blas_threads(NCORES):
        tr = oracle_trainer(name, rows)
        x, t = CFG.synthetic_bunch(name, 42, rows)
        t0 = time.time()
        for _ in range(max(warmup, 1)):
            tr.train_step(x, t)
            if time.time() - t0 > budget_s / 3:
                break
        t0 = time.time()
        done = 0
        for _ in range(steps):
            tr.train_step(x, t)
            done += 1
            if time.time() - t0 > budget_s:
                break
        dt = time.time() - t0
    return done * rows / dt, dt / done, done, rows


def cpu_sample_text(name, done, rows):
    full = CFG.CONFIGS[name]["bunch"]
    what = "full bunch" if rows == full else "bounded sample: %d of the bunch's %d rows per step" % (rows, full)
    return ("%d full train steps of %s rows (%s) on the host, numpy/OpenBLAS port of the reference algorithm -- a FASTER "
            "baseline than the reference's own binary, which measured 2848 samples/s on C2 with 8 threads (BASELINE.md)"
            % (done, rows, what))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, s_per_step, done, rows = time_oracle(args.config, args.steps, args.warmup, budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "warmup": args.warmup, "ms_per_step": 1e3 * s_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": CFG.workload_string(args.config)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores_used(), "kind": "port",
                         "sample": cpu_sample_text(args.config, done, rows)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------- parity phase
def parity_phase(args, ann, tr, ctx, dist, rank, world):
    """Every rank: same seeded weights, zero optimizer state, count 0, then k train steps on fixed bunches
    (seeded by rank and step) through the data-parallel path that was just timed.  Rank 0: the oracle's k steps
    on the concatenated global bunch from the same weights; compares the weight MOVEMENT and the last loss."""
    name = args.config
    k, pb = int(os.environ.get("B200_PARITY_STEPS", PARITY_STEPS[name])), PARITY_BUNCH[name]
    names = tr.weight_names()
    rng = np.random.RandomState(4321)          # same stream on every rank -> identical replicas
    w0 = {}
    for n in names:
        shape = tr._dims(n)
        w0[n] = (rng.uniform(-1, 1, size=shape) / np.sqrt(shape[0] + shape[1])).astype(np.float32)
        tr.set_weights(n, w0[n])
        tr.set_weights(n, np.zeros(shape, np.float32), which=2)
    tr.set_count(0)
    if dist is not None:
        dist.barrier()
    last = 0.0
    for s in range(k):
        x, t = CFG.synthetic_bunch(name, 9000 + 1000 * s + rank, pb)
        last, _ = tr.train_step(x, t, bunch_size=pb)
    ctx.sync()
    if dist is not None:
        dist.barrier()
    # replicas must hold identical weights: compare a float64 checksum over every tensor
    w1 = {n: tr.weights(n) for n in names}
    chk = float(sum(np.abs(w1[n].astype(np.float64)).sum() for n in names))
    spread = 0.0
    if dist is not None:
        import torch
        tt = torch.tensor([chk, last], dtype=torch.float64)
        lst = [torch.zeros_like(tt) for _ in range(world)]
        dist.all_gather(lst, tt)
        spread = max(abs(v[0].item() - chk) for v in lst)
        last = float(np.mean([v[1].item() for v in lst]))
    if rank != 0:
        return None
    t0 = time.time()
    with blas_threads(NCORES):
        ref = oracle_trainer(name, world * pb, weights=w0)
        for s in range(k):
            xs, ts = zip(*[CFG.synthetic_bunch(name, 9000 + 1000 * s + r, pb) for r in range(world)])
            ref_last, _ = ref.train_step(np.concatenate(xs), np.concatenate(ts), bunch_size=world * pb)
    num = den = 0.0
    worst = ("", 0.0)
    for n in names:
        dg = w1[n].astype(np.float64) - w0[n]
        dr = ref.weights[n].astype(np.float64).reshape(w0[n].shape) - w0[n]
        e, d = float(((dg - dr) ** 2).sum()), float((dr ** 2).sum())
        num += e
        den += d
        if d > 0 and np.sqrt(e / d) > worst[1]:
            worst = (n, float(np.sqrt(e / d)))
    rel = float(np.sqrt(num / max(den, 1e-300)))
    relu_like = name in ("C2", "C4")
    tol = (3e-2 if relu_like else 5e-3) if args.math == "tf32" else (2e-3 if relu_like else 1e-4)
    return {"rel_l2_w": rel, "worst_tensor": worst[0], "worst_rel_l2": worst[1], "loss_abs": abs(last - ref_last),
            "loss_gpu": last, "loss_oracle": ref_last, "steps": k, "global_bunch": world * pb, "mode": args.math,
            "n_gpus": world, "replica_checksum_spread": spread, "tolerance": tol,
            "ok": bool(rel < tol and abs(last - ref_last) < 2e-3 * max(1.0, abs(ref_last)) and spread == 0.0),
            "what": "rel-L2 of the weight movement over k steps (all tensors) vs the CPU oracle on the concatenated "
                    "global bunch, from identical seeded weights; ReLU/max-pool nets: units within rounding of a gate "
                    "flip, hence the looser tolerance (tests/test_gpu_fullsize.py)",
            "oracle_seconds": time.time() - t0}


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

    name = args.config
    cfg = CFG.CONFIGS[name]
    BUNCH, in_sz, out_sz = cfg["bunch"], cfg["nin"], cfg["nout"]
    ctx = ann.get_context(local_rank)
    mode = ann.MATH_FP32 if args.math == "fp32" else ann.MATH_TF32
    ctx.set_math_mode(mode)
    tr = CFG.build_trainer(ann, name, ctx=ctx)
    tr.randomize_weights(random=ann.random(1234), inf=-1, sup=1, use_fanin=True, use_fanout=True)
    if world > 1:
        init_data_parallel(tr, dist)

    # pinned host bunches (distinct per rank and rotated per step)
    nb = 4
    hx = C.c_void_p()
    ht = C.c_void_p()
    check(lib.b200_host_alloc(C.byref(hx), C.c_size_t(nb * BUNCH * in_sz * 4)))
    check(lib.b200_host_alloc(C.byref(ht), C.c_size_t(nb * BUNCH * out_sz * 4)))
    px = np.ctypeslib.as_array(C.cast(hx, C.POINTER(C.c_float)), shape=(nb, BUNCH, in_sz))
    pt = np.ctypeslib.as_array(C.cast(ht, C.POINTER(C.c_float)), shape=(nb, BUNCH, out_sz))
    for i in range(nb):
        px[i], pt[i] = CFG.synthetic_bunch(name, 42 + 100 * rank + i)
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
    W = max(args.warmup, 3)
    tr.stage(px[0], pt[0], BUNCH)
    for _ in range(W):
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
    for k in range(max(W, 6)):
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

    parity = None
    if not args.no_parity:
        parity = parity_phase(args, ann, tr, ctx, dist, rank, world)

    if rank != 0:
        if dist is not None:
            dist.barrier()
        return

    pk = peaks()
    value = world * BUNCH * K / (dev_ms / 1e3)
    flops = CFG.step_flops(name)
    # tensor peak for the compute type: TF32 runs at half the bf16 rate on the tcgen05 pipe;
    # the denominator is the MEASURED bf16 cuBLAS number / 2.  fp32 (FFMA) mode has no tensor peak.
    tf32_peak = pk["bf16_sustained"] / 2.0
    step_tflops = flops / (dev_ms / K / 1e3) / 1e12

    roof = kernel_roofline(name, ann, ctx, lib, check, C, pk, args.math)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "tf32" if args.math == "tf32" else "f32", "data": "synthetic",
        "config": {"workload": CFG.workload_string(name),
                   "global_bunch": world * BUNCH, "parallelism": "dp%d" % world, "math": args.math,
                   "l2": "flushed between steps (256 MiB memset outside the per-step event bracket)",
                   "timing": "sum of per-step CUDA-event durations on the launching stream, max over ranks"},
        "clocks": clocks,
        "e2e": {"value": world * BUNCH * K / (e2e_ms / 1e3), "unit": UNIT,
                "h2d_bytes_per_step": BUNCH * (in_sz + out_sz) * 4, "d2h_bytes_per_step": 8,
                "wall_value": world * BUNCH * K / e2e_wall, "last_loss": last_loss},
        "gpu_launches": int(launches),
        "roofline": roof,
        "parity": parity,
        "step_tflops": step_tflops, "step_frac_of_tf32_peak": step_tflops / tf32_peak,
        "tf32_peak_tflops": tf32_peak, "peak_source": pk["source"] + "; tf32 = bf16_sustained/2",
        "hot_l2_value": world * BUNCH * K / (hot_ms / 1e3),
        "per_step_ms_minmax": [min(per_step), max(per_step)],
    }
    if world == 1:
        cpu_val, cpu_s, cpu_done, cpu_rows = time_oracle(name, 40, 1, budget_s=15.0)
        line["cpu_baseline"] = {"value": cpu_val, "unit": UNIT, "cores": cores_used(), "kind": "port",
                                "sample": cpu_sample_text(name, cpu_done, cpu_rows)}
    print(json.dumps(line))
    if dist is not None:
        dist.barrier()


def _time_launches(ctx, lib, check, C, launch, nsets, reps):
    e0, e1 = C.c_void_p(), C.c_void_p()
    check(lib.b200_event_create(C.byref(e0)))
    check(lib.b200_event_create(C.byref(e1)))
    for i in range(max(2 * nsets, 4)):
        launch(i)
    ctx.sync()
    check(lib.b200_event_record(ctx.h, e0))
    for i in range(reps):
        launch(i)
    check(lib.b200_event_record(ctx.h, e1))
    ms = C.c_float()
    check(lib.b200_event_elapsed_ms(e0, e1, C.byref(ms)))
    return ms.value / reps / 1e3


def _traffic(tag):
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            return json.load(open(tpath)).get(tag)
        except Exception:
            return None
    return None


def kernel_roofline(name, ann, ctx, lib, check, C, pk, math):
    """The dominant kernel of the workload, timed alone: `reps` launches back to back inside one CUDA-event
    bracket on the launching stream, rotating over operand sets that together exceed the 126 MB L2 so that no
    launch finds its operands cached; achieved = algorithmic FLOPs (or bytes) per launch / mean launch time.
      C1 / C2 / C3: the largest dense forward contraction (+ bias + activation epilogue)   -> tensor roofline
      C4: the second convolution's forward pass (+ bias + ReLU)                             -> tensor roofline
      C5: the fused log_softmax + MCCE + gradient pass over 4096 x 10000                    -> HBM roofline"""
    from april_ann_b200.ops import DeviceArray
    rng = np.random.RandomState(0)
    I = C.c_int
    if math == "tf32":
        tpeak = pk["bf16_burst"] / 2.0
        tnote = "tf32 tensor peak = measured bf16 burst / 2 (%s)" % pk["source"]
    else:
        tpeak = 75.0
        tnote = "fp32 FFMA mode: nominal 75 TFLOP/s CUDA-core peak (no measured figure)"

    def dev(*shape, lo=-1.0, hi=1.0):
        return DeviceArray.from_numpy(ctx, rng.uniform(lo, hi, shape).astype(np.float32))

    if name == "C5":
        M, Cc = 4096, 10000
        sets = []
        for _ in range(2):
            t = DeviceArray(ctx, (M, Cc))
            t.zero()
            sets.append((dev(M, Cc, lo=-4, hi=4), t, DeviceArray(ctx, (M, Cc)), DeviceArray(ctx, (M,)), DeviceArray(ctx, (M, Cc))))

        def launch(i):
            z, t, lp, rows, g = sets[i % len(sets)]
            check(lib.b200_log_softmax_mcce_fused(ctx.h, I(M), I(Cc), z.ptr, t.ptr, lp.ptr, rows.ptr, g.ptr))
        tsec = _time_launches(ctx, lib, check, C, launch, len(sets), 20)
        nbytes = 16.0 * M * Cc
        ach = nbytes / tsec / 1e9
        return {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                "traffic": _traffic("C5"), "kernel": "lsm_mcce_wide_kernel: log_softmax + MCCE rows + gradient, %d x %d, 16 B per (row, class)" % (M, Cc),
                "launch_us": tsec * 1e6, "note": "HBM peak = measured copy bandwidth (%s); 20 launches over 2 operand sets (1.3 GB > L2)" % pk["source"]}
    if name == "C4":
        B, Cin, H, Wd, n, k = 512, 16, 12, 12, 32, 5
        sets = [(dev(B, Cin, H, Wd), dev(n, Cin * k * k, lo=-.1, hi=.1), DeviceArray(ctx, (B, n, H - k + 1, Wd - k + 1))) for _ in range(16)]
        bias = dev(n)

        def launch(i):
            x, w, y = sets[i % len(sets)]
            check(lib.b200_conv2d_fwd(ctx.h, I(B), I(Cin), I(H), I(Wd), I(n), I(k), I(k), I(1), I(1), x.ptr, w.ptr, bias.ptr, I(3), y.ptr))
        tsec = _time_launches(ctx, lib, check, C, launch, len(sets), 64)
        fl = 2.0 * B * (H - k + 1) * (Wd - k + 1) * n * Cin * k * k
        ach = fl / tsec / 1e12
        return {"bound": "tensor", "achieved": ach, "peak": tpeak, "unit": "TFLOP/s", "frac": ach / tpeak, "traffic": _traffic("C4"),
                "kernel": "conv2d_fwd [512,16,12,12] * 5x5x32 (+bias+relu): a 32768 x 32 x 400 contraction, %s" % math,
                "launch_us": tsec * 1e6, "note": tnote + "; 64 launches over 16 operand sets"}
    M, N, K, act, nsets, reps = {"C1": (32, 256, 256, 2, 8, 64), "C2": (1024, 2048, 2048, 3, 8, 64),
                                 "C3": (8192, 4096, 4096, 2, 2, 10)}[name]
    sets = [(dev(M, K), dev(N, K, lo=-.05, hi=.05), DeviceArray(ctx, (M, N))) for _ in range(nsets)]
    b = dev(N, lo=-.1, hi=.1)

    def launch(i):
        X, Wm, Y = sets[i % nsets]
        check(lib.b200_linear_fwd(ctx.h, I(M), I(N), I(K), X.ptr, I(K), Wm.ptr, I(K), b.ptr, I(act), Y.ptr, I(N)))
    tsec = _time_launches(ctx, lib, check, C, launch, nsets, reps)
    ach = 2.0 * M * N * K / tsec / 1e12
    out = {"bound": "tensor", "achieved": ach, "peak": tpeak, "unit": "TFLOP/s", "frac": ach / tpeak, "traffic": _traffic(name),
           "kernel": "linear_fwd M=%d N=%d K=%d (+bias+%s epilogue), %s" % (M, N, K, "relu" if act == 3 else "tanh", math),
           "launch_us": tsec * 1e6,
           "note": tnote + "; %d launches back to back over %d operand sets (%d MiB in all)" % (
               reps, nsets, nsets * 4 * (M * K + N * K + M * N) >> 20)}
    if name == "C1":
        out["note"] += "; a 32-row bunch is launch/latency bound, not tensor bound"
    return out


def main():
    if ARGS.impl == "reference":
        run_reference(ARGS)
    else:
        run_b200(ARGS)


if __name__ == "__main__":
    main()
