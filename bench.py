#!/usr/bin/env python
"""bench.py -- images/sec of one AlexNet-lite train step (BASELINE.json metric, config 2).

    python bench.py --gpus 1 --steps 20 --warmup 5            # our arm, one JSON line
    torchrun --nproc-per-node N ... bench.py --gpus N ...      # data parallel, weak scaling
    python bench.py --impl reference ...                       # the reference's CPU code

A step = forward + softmax/cross-entropy + backward (all three conv gradients, image gradient
included, as the reference computes it) [+ gradient all-reduce when N>1] + SGD, on a batch of
B=256 synthetic 3x224x224 fp32 images per GPU, random-init (reference-seed) weights.
`value` is timed with CUDA events on the launching stream with inputs resident in HBM; `e2e`
goes through the host-buffer C-ABI call (cnn_net_train_step_host: H2D of the batch from pinned
memory, the step, D2H of loss + probabilities, every step).
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "images/sec (train step, 224x224x3)"
UNIT = "images/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, "fallback"


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------- reference / cpu baseline
def _ref_worker(args):
    """One UNMODIFIED-reference process: AlexNet (alexnet.cpp) train steps at batch b."""
    b, steps, warm, seed = args
    from oracle import ref, port
    from cnn_b200.nets import alexnet_lite
    from cnn_b200.synth import synth_images, synth_labels
    x, lab = synth_images(b, seed=seed), synth_labels(b)
    init = np.fromfile(os.path.join(ROOT, "tests", "golden", "alexnet_init.model"), np.float32)
    if ref.available():
        net, kind = ref.Net(), "reference"
        net.set_params(init)
        step = lambda: net.train_step(x, lab, 1e-3)
    else:
        net, kind = port.Net(alexnet_lite(3), b, 3, 224, 224), "port"
        net.set_params(init)
        step = lambda: net.train_step(x, lab, 1e-3)
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    return kind, time.perf_counter() - t0


def cpu_reference_rate(procs, b, steps, warm):
    """images/s of `procs` independent single-threaded reference processes (the reference has
    no threads, SIMD or BLAS; replicas are the only way it can use more cores)."""
    import multiprocessing as mp
    if procs == 1:
        kind, dt = _ref_worker((b, steps, warm, 1234))
        return kind, b * steps / dt, dt
    with mp.get_context("spawn").Pool(procs) as pool:
        t0 = time.perf_counter()
        res = pool.map(_ref_worker, [(b, steps, warm, 1234 + i) for i in range(procs)])
        wall = time.perf_counter() - t0
    kind = res[0][0]
    rate = sum(b * steps / dt for _, dt in res)
    return kind, rate, wall


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, a.ref_procs if a.ref_procs > 0 else cores))
    b = 4  # the reference's own train batch (cnn.cpp:36); each step a bounded sample of the workload
    steps = max(1, min(a.steps, 8))
    warm = 1
    t0 = time.perf_counter()
    kind, rate, _ = cpu_reference_rate(procs, b, steps, warm)
    _, single, _ = (kind, rate, 0) if procs == 1 else cpu_reference_rate(1, b, min(steps, 4), 1)
    wall = time.perf_counter() - t0
    line = {
        "impl": "reference", "metric": METRIC, "value": round(rate, 3), "unit": UNIT, "n_gpus": a.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": round(1000.0 * b * procs / rate, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": "alexnet_lite train step, 3x224x224 fp32, CPU reference",
                                        "batch_per_process": b, "processes": procs},
        "cpu_baseline": {"value": round(rate, 3), "unit": UNIT, "cores": procs, "kind": kind,
                         "single_thread_value": round(single, 3),
                         "sample": f"{procs} independent single-threaded processes x {steps} steps x batch {b}"
                                   f" (host has {cores} cores; wall {wall:.1f}s)"},
        "e2e": {"value": round(rate, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    OUT.emit(json.dumps(line))


# ------------------------------------------------------------------------------ our arm
def op_breakdown(ctx, net_spec, B, reps=5):
    """CUDA-event time of every operator of one train step, launched eagerly through the same
    C-ABI entry points the engine uses, with the algorithmic bytes / flops of each."""
    import torch
    from cnn_b200 import nets
    C, H, W = 3, 224, 224
    rows = []
    x = torch.rand(B, C, H, W, device=ctx.device)

    def timed(fn):
        with torch.cuda.stream(ctx.stream):
            fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ctx.stream)
            for _ in range(reps):
                fn()
            e1.record(ctx.stream)
        e1.synchronize()
        return e0.elapsed_time(e1) / reps * 1e-3

    cur = x
    saved = []
    skip_next = False
    for li, (t, a, b, c, d) in enumerate(net_spec):
        if skip_next:
            skip_next = False
            continue
        if t == nets.CONV:
            w = torch.randn(b, a, c, c, device=ctx.device) / 10
            bias = torch.zeros(b, device=ctx.device)
            y = ctx.conv2d_forward(cur, w, bias, d)
            fl = 2.0 * y.numel() * a * c * c
            xin = cur
            rows.append((f"conv{li}.fwd", timed(lambda: ctx.conv2d_forward(xin, w, bias, d)),
                         4.0 * (xin.numel() + y.numel()), fl))
            saved.append(("conv", li, xin, w, y, d, fl))
            cur = y
        elif t == nets.RELU and li + 1 < len(net_spec) and net_spec[li + 1][0] == nets.POOL and net_spec[li + 1][2] >= net_spec[li + 1][1]:
            # the engine runs ReLU + MaxPool as one kernel each way (both layers' outputs are written)
            xin = cur
            pk, ps = net_spec[li + 1][1], net_spec[li + 1][2]
            yr, yp, mask = ctx.relu_maxpool_forward(xin, pk, ps)
            rows.append((f"relu{li}+pool{li + 1}.fwd", timed(lambda: ctx.relu_maxpool_forward(xin, pk, ps)),
                         8.0 * xin.numel() + 8.0 * yp.numel(), 0.0))
            saved.append(("relupool", li, xin.shape, mask, pk, ps, yp))
            cur = yp
            skip_next = True
        elif t == nets.RELU:
            xin = cur
            y = ctx.relu_forward(xin)
            rows.append((f"relu{li}.fwd", timed(lambda: ctx.relu_forward(xin)), 8.0 * xin.numel(), 0.0))
            saved.append(("relu", li, y))
            cur = y
        elif t == nets.POOL:
            xin = cur
            y, mask = ctx.maxpool_forward(xin, a, b)
            rows.append((f"pool{li}.fwd", timed(lambda: ctx.maxpool_forward(xin, a, b)),
                         4.0 * xin.numel() + 8.0 * y.numel(), 0.0))
            saved.append(("pool", li, xin.shape, mask, a, b, y))
            cur = y
        elif t == nets.LINEAR:
            xin = cur.reshape(B, -1)
            w = torch.randn(a, b, device=ctx.device) / 10
            bias = torch.zeros(b, device=ctx.device)
            y = ctx.linear_forward(xin, w, bias)
            rows.append((f"linear{li}.fwd", timed(lambda: ctx.linear_forward(xin, w, bias)),
                         4.0 * (xin.numel() + w.numel()), 2.0 * B * a * b))
            saved.append(("linear", li, xin, w, y))
            cur = y
        elif t == nets.BN:
            xin = cur
            g, bt = torch.ones(a, device=ctx.device), torch.zeros(a, device=ctx.device)
            mm, mv = torch.zeros(a, device=ctx.device), torch.zeros(a, device=ctx.device)
            r = ctx.bn_forward_train(xin, g, bt, mm, mv)
            rows.append((f"bn{li}.fwd", timed(lambda: ctx.bn_forward_train(xin, g, bt, mm, mv)),
                         20.0 * xin.numel(), 0.0))
            saved.append(("bn", li, xin, r, g))
            cur = r["y"]
    for item in reversed(saved):
        kind, li = item[0], item[1]
        if kind == "conv":
            _, _, xin, w, y, d, fl = item
            delta = torch.randn_like(y)
            L, h = ctx.L, ctx._h
            Bc, Cin, Hc, Wc = xin.shape
            Cout, _, k, _ = w.shape
            dw, db, dx = torch.empty_like(w), torch.empty(Cout, device=ctx.device), torch.empty_like(xin)
            import ctypes as CT
            P = lambda tt: CT.c_void_p(tt.data_ptr())
            rows.append((f"conv{li}.wgrad", timed(lambda: L.cnn_conv2d_backward_weights(
                h, P(xin), P(delta), P(dw), P(db), Bc, Cin, Hc, Wc, Cout, k, d, 1.0 / B)),
                4.0 * (xin.numel() + y.numel()), fl))
            rows.append((f"conv{li}.dgrad", timed(lambda: L.cnn_conv2d_backward_data(
                h, P(w), P(delta), P(dx), Bc, Cin, Hc, Wc, Cout, k, d)),
                4.0 * (xin.numel() + y.numel()), fl))
        elif kind == "relu":
            y = item[2]
            delta = torch.randn_like(y)
            rows.append((f"relu{li}.bwd", timed(lambda: ctx.relu_backward(delta, y)), 12.0 * y.numel(), 0.0))
        elif kind == "relupool":
            _, _, shp, mask, a, b, yp = item
            delta = torch.randn_like(yp)
            n_in = int(np.prod(shp))
            rows.append((f"relu{li}+pool{li + 1}.bwd", timed(lambda: ctx.maxpool_relu_backward(delta, mask, yp, shp, a, b)),
                         4.0 * n_in + 12.0 * yp.numel(), 0.0))
        elif kind == "pool":
            _, _, shp, mask, a, b, y = item
            delta = torch.randn_like(y)
            n_in = int(np.prod(shp))
            rows.append((f"pool{li}.bwd", timed(lambda: ctx.maxpool_backward(delta, mask, shp, a, b)),
                         4.0 * n_in + 8.0 * y.numel(), 0.0))
        elif kind == "linear":
            _, _, xin, w, y = item
            delta = torch.randn_like(y)
            rows.append((f"linear{li}.bwd", timed(lambda: ctx.linear_backward(xin, w, delta)),
                         4.0 * (2 * xin.numel() + 2 * w.numel()), 6.0 * B * w.numel()))
        elif kind == "bn":
            _, _, xin, r, g = item
            delta = torch.randn_like(xin)
            rows.append((f"bn{li}.bwd", timed(lambda: ctx.bn_backward(delta, xin, r["xhat"], g, r["mean"], r["var"])),
                         28.0 * xin.numel(), 0.0))
    return rows


def run_ours(a):
    import torch
    import torch.distributed as dist
    from cnn_b200 import nets
    from cnn_b200.api import Context, Net
    from cnn_b200.synth import synth_images, synth_labels

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == a.gpus, f"--gpus {a.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = a.batch
    spec = nets.alexnet_lite(3) if a.net == "alexnet_lite" else nets.vgg_style(3)
    ctx = Context(local)
    if a.conv_algo != "auto":
        from cnn_b200 import api
        ctx.set_conv_algo({"simt": api.CONV_SIMT, "tcgen05": api.CONV_TCGEN05}[a.conv_algo])
    net = Net(ctx, spec, B)
    init = np.fromfile(os.path.join(ROOT, "tests", "golden", "alexnet_init.model"), np.float32)
    if a.net == "alexnet_lite":
        net.set_params(init)
    else:
        rng = np.random.default_rng(0)
        net.set_params((rng.standard_normal(net.n_params) * 0.02).astype(np.float32))
    # two resident input batches (> L2 each: 154 MB at B=256) alternate between steps
    xs = [ctx.to_device(synth_images(B, seed=1234 + i, first_image=rank * B)) for i in range(2)]
    lab = ctx.to_device(synth_labels(B, 3, first_image=rank * B), torch.int32)
    from cnn_b200.dist import NetEngine, dp_train_step, init_native_dist
    native = False
    if world > 1 and not a.torch_allreduce:
        # the library's own communicator: the slab all-reduce then sits inside the step's CUDA graph
        try:
            init_native_dist(ctx)
            native = True
        except Exception as e:  # keep measuring through torch.distributed, and say so
            print(f"[bench] library NCCL unavailable ({e}); all-reduce through torch.distributed", file=sys.stderr)
        flag = torch.tensor([1 if native else 0], device=ctx.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)      # all ranks take the same path
        native = bool(flag.item())
    engine = NetEngine(net, native_dist=native)
    slab = net.grad_slab()
    scale = 1.0 / (B * world)
    lr = 1e-3

    def step(i):
        if world == 1:
            net.train_step(xs[i & 1], lab, lr, grad_scale=scale, do_update=True)
        else:  # fwd+bwd graph, ONE NCCL all-reduce of the slab (gradients + loss tail), replicated SGD
            dp_train_step(engine, xs[i & 1], lab, lr, B * world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(a.warmup, 3)):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ctx.stream)
    for i in range(a.steps):
        step(i)
    e1.record(ctx.stream)
    barrier()
    launches = ctx.launches - l0
    ms = e0.elapsed_time(e1)
    clocks = None
    loss = float(net.loss_from_slab(B * world))
    t = torch.tensor([ms], device=ctx.device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ms_step = ms / a.steps
    value = B * world / (ms_step * 1e-3)
    # K steps of < 1 ms are over before nvidia-smi (100 ms period) has sampled twice: the same steps keep
    # running, untimed, for another ~0.4 s so that the clock / throttle record describes this load (the
    # count comes from the rank-maximum step time, so every rank runs the same number of collectives)
    n_obs = max(20, min(2000, int(400.0 / max(ms_step, 1e-3))))
    for i in range(n_obs):
        step(a.steps + i)
    barrier()
    if rank == 0:
        clocks = sampler.stop()
        clocks["window"] = f"timed region + {n_obs} more identical steps (untimed)"

    # ---- end to end through the host-buffer C-ABI call (pinned host memory) ----------
    hx = [torch.from_numpy(synth_images(B, seed=4321 + i, first_image=rank * B)).pin_memory() for i in range(2)]
    hl = torch.from_numpy(synth_labels(B, 3, first_image=rank * B)).pin_memory()
    hp = torch.empty(B, net.classes).pin_memory()
    e2e_steps = max(3, min(a.steps, 10))
    e2e_extra = {}
    if world == 1:
        def timed_host_loop(run):
            barrier()
            t0 = time.perf_counter()
            e0.record(ctx.stream)
            run()
            e1.record(ctx.stream)
            barrier()
            return max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3) / e2e_steps

        def blocking():
            for i in range(e2e_steps):
                net.train_step_host(hx[i & 1], hl, lr, hp)

        def piped(src):
            # every step: H2D of its batch (copy stream, overlapped with the previous step), the step,
            # D2H of loss + probabilities read by the host in wait_host
            def run():
                net.submit_host(src[0], hl, lr)
                for i in range(1, e2e_steps):
                    net.submit_host(src[i & 1], hl, lr)
                    net.wait_host(hp)
                net.wait_host(hp)
            return run

        h8 = [torch.from_numpy(np.random.default_rng(5 + i).integers(0, 256, (B, 224, 224, 3), dtype=np.uint8)).pin_memory()
              for i in range(2)]
        for fn in (blocking, piped(hx), piped(h8)):   # warm-up: staging buffers, graphs for both slots
            fn()
        blocking_ms = timed_host_loop(blocking)
        e2e_ms = timed_host_loop(piped(hx))
        u8_ms = timed_host_loop(piped(h8))
        e2e_extra = {
            "call": "cnn_net_train_step_host_submit/_wait, fp32 host images, depth-2 pipeline",
            "blocking_call": {"value": round(B / (blocking_ms * 1e-3), 1), "ms_per_step": round(blocking_ms, 4),
                              "call": "cnn_net_train_step_host"},
            "u8_images": {"value": round(B / (u8_ms * 1e-3), 1), "ms_per_step": round(u8_ms, 4),
                          "h2d_bytes_per_step": B * 3 * 224 * 224 + B * 4,
                          "call": "cnn_net_train_step_host_submit_u8/_wait (loader bytes, read_from_opencv_mat on device)"},
        }
    else:
        # data parallel: the same depth-2 pipeline built from torch streams -- the H2D of batch i+1 runs on a
        # copy stream while step i (fwd+bwd graph, NCCL all-reduce, SGD) runs on the compute stream; every
        # step's loss + probabilities are read back to pinned host memory inside the timed region
        copy_stream = torch.cuda.Stream(device=ctx.device)
        labs = [lab, torch.empty_like(lab)]
        hloss = [torch.empty(1).pin_memory() for _ in range(2)]
        hps = [torch.empty(B, net.classes).pin_memory() for _ in range(2)]
        copied = [torch.cuda.Event() for _ in range(2)]
        stepped = [torch.cuda.Event() for _ in range(2)]

        def submit(i):
            sl = i & 1
            with torch.cuda.stream(copy_stream):
                if i >= 2:
                    copy_stream.wait_event(stepped[sl])
                xs[sl].copy_(hx[sl], non_blocking=True)
                labs[sl].copy_(hl, non_blocking=True)
                copied[sl].record(copy_stream)
            ctx.stream.wait_event(copied[sl])
            dp_train_step(engine, xs[sl], labs[sl], lr, B * world)
            with torch.cuda.stream(ctx.stream):
                hloss[sl].copy_(slab[-1:], non_blocking=True)
                hps[sl].copy_(net.probs(), non_blocking=True)
                stepped[sl].record(ctx.stream)

        def piped_dp(steps):
            submit(0)
            for i in range(1, steps):
                submit(i)
                stepped[(i - 1) & 1].synchronize()
            stepped[(steps - 1) & 1].synchronize()

        piped_dp(4)
        barrier()
        t0 = time.perf_counter()
        piped_dp(e2e_steps)
        barrier()
        tt = torch.tensor([(time.perf_counter() - t0) * 1e3], device=ctx.device, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt.item()) / e2e_steps
        e2e_extra = {"call": "depth-2 pipeline: pinned fp32 host batch -> copy stream H2D -> dp_train_step (graph + one NCCL "
                             "all-reduce + SGD) -> D2H loss/probabilities, per rank"}
    e2e_value = B * world / (e2e_ms * 1e-3)
    h2d = B * 3 * 224 * 224 * 4 + B * 4
    d2h = 4 + B * net.classes * 4

    line = None
    if rank == 0:
        hbm, tc, which = peaks()
        rows = op_breakdown(ctx, spec, B) if not a.no_breakdown else []
        roof, breakdown = None, None
        if rows:
            tot = sum(r[1] for r in rows)
            breakdown = {n: {"us": round(t_ * 1e6, 1), "share": round(t_ / tot, 4),
                             "GBps": round(by / t_ / 1e9, 1), "TFLOPs": round(fl / t_ / 1e12, 2)}
                         for n, t_, by, fl in rows}
            n, t_, by, fl = max(rows, key=lambda r: r[1])
            ach = by / t_ / 1e9
            traffic = None
            try:
                with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                    traffic = json.load(f).get(n)
            except Exception:
                pass
            roof = {"kernel": n, "bound": "hbm", "achieved": round(ach, 1), "peak": hbm, "unit": "GB/s",
                    "frac": round(ach / hbm, 4), "traffic": traffic, "peak_source": which,
                    "algorithmic_bytes": by, "launch_us": round(t_ * 1e6, 1),
                    "flops": fl, "tflops": round(fl / t_ / 1e12, 2)}
        cpu = None
        if world == 1 and not a.no_cpu:
            kind, rate, dt = cpu_reference_rate(1, 4, a.cpu_steps, 1)
            cpu = {"value": round(rate, 3), "unit": UNIT, "cores": 1, "kind": kind,
                   "sample": f"{a.cpu_steps} train steps at batch 4 (the reference's own batch, cnn.cpp:36), "
                             f"single thread, {dt:.1f}s; host has {os.cpu_count()} cores"}
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": round(ms_step, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{a.net} train step (fwd+xent+bwd incl. image grad+SGD), "
                                   f"batch {B}/GPU x {world} GPU, 3x224x224 fp32, reference-seed init",
                       "global_batch": B * world, "parallelism": f"dp{world}", "conv_algo": a.conv_algo,
                       "allreduce": ("none" if world == 1 else "ncclAllReduce issued by the library inside the step graph"
                                     if native else "torch.distributed all_reduce between graph and SGD"),
                       "l2": "two alternating resident input batches of 154 MB each and ~1.5 GB of "
                             "activations per step exceed the 126 MB L2",
                       "cuda_graph": True},
            "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": round(e2e_ms, 4), "steps": e2e_steps, **e2e_extra},
            "gpu_launches": int(launches),
            "clocks": clocks, "loss_after": loss,
            "roofline": roof, "cpu_baseline": cpu, "breakdown": breakdown,
        }
    net.close()
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line:
        OUT.emit(json.dumps(line))


class OneLineStdout:
    """The contract is ONE JSON line on stdout.  Libraries loaded by this process write there too (NCCL prints
    its version banner on communicator creation), so file descriptor 1 points at stderr for the whole run
    and the line goes out through the saved descriptor."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, text):
        sys.stdout.flush()
        os.write(self.saved, (text + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


OUT = None


def main():
    global OUT
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="images per GPU")
    ap.add_argument("--net", default="alexnet_lite", choices=["alexnet_lite", "vgg_style"])
    ap.add_argument("--conv-algo", default="auto", choices=["auto", "simt", "tcgen05"])
    ap.add_argument("--cpu-steps", type=int, default=40)
    ap.add_argument("--ref-procs", type=int, default=0, help="reference arm processes (0 = all cores)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--torch-allreduce", action="store_true", help="N>1: all-reduce through torch.distributed")
    ap.add_argument("--no-breakdown", action="store_true")
    a = ap.parse_args()
    with OneLineStdout() as OUT:
        if a.impl == "reference":
            run_reference(a)
        else:
            run_ours(a)


if __name__ == "__main__":
    main()
