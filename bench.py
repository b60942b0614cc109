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
    # --steps / --warmup are honoured; a B=4 CPU step is ~0.1 s, so even 50 steps end within seconds
    steps = max(1, min(a.steps, 200))
    warm = max(0, min(a.warmup, 20))
    t0 = time.perf_counter()
    kind, rate, _ = cpu_reference_rate(procs, b, steps, warm)
    _, single, _ = (kind, rate, 0) if procs == 1 else cpu_reference_rate(1, b, min(steps, 8), 1)
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
def layer_costs(spec, B, C=3, H=224, W=224):
    """Algorithmic FLOPs and compulsory fp32 bytes per (layer, pass) -- SURVEY 8(d): conv pass
    2*B*OH*OW*Cout*Cin*k^2 FLOP, bytes 4*(X + Y) (wgrad: X + delta, dgrad: delta + dX); Linear pass 2*B*in*out;
    ReLU fwd 8 / bwd 12 B per element, MaxPool fwd 4*in + 8*out, bwd 4*in + 8*out, BN fwd 16 / bwd 16."""
    from cnn_b200 import nets
    out = {}
    c, h, w = C, H, W
    for li, ((t, a, b, kk, d), (oc, oh, ow)) in enumerate(zip(spec, nets.shapes(spec, C, H, W))):
        nin, nout = B * c * h * w, B * oc * oh * ow
        if t == nets.CONV:
            fl = 2.0 * nout * a * kk * kk
            by = 4.0 * (nin + nout)
            out[(li, "f")] = (fl, by, "conv")
            out[(li, "b")] = (2 * fl, 2 * by, "conv")   # weight + input gradient
        elif t == nets.LINEAR:
            fl = 2.0 * B * a * b
            out[(li, "f")] = (fl, 4.0 * (nin + a * b + nout), "linear")
            out[(li, "b")] = (2 * fl, 4.0 * (2 * nin + 2 * a * b + nout), "linear")
        elif t == nets.RELU:
            out[(li, "f")] = (0.0, 8.0 * nin, "relu")
            out[(li, "b")] = (0.0, 12.0 * nin, "relu")
        elif t == nets.POOL:
            out[(li, "f")] = (0.0, 4.0 * nin + 8.0 * nout, "pool")
            out[(li, "b")] = (0.0, 4.0 * nin + 8.0 * nout, "pool")
        elif t == nets.BN:
            out[(li, "f")] = (0.0, 16.0 * nin, "bn")
            out[(li, "b")] = (0.0, 16.0 * nin, "bn")
        c, h, w = oc, oh, ow
    return out


def profile_step(ctx, net, x, lab, lr, scale, do_update, reps=3):
    """Per-launch CUDA-event times (cnn_prof_*) of eager train steps -- the same kernels the graph replays."""
    net.use_graph(False)
    net.train_step(x, lab, lr, grad_scale=scale, do_update=do_update)   # warm (allocations, plans)
    acc = None
    for _ in range(reps):
        rows = ctx.profile(lambda: net.train_step(x, lab, lr, grad_scale=scale, do_update=do_update))
        if acc is None:
            acc = [[n, t] for n, t in rows]
        else:
            for r, (n, t) in zip(acc, rows):
                r[1] += t
    net.use_graph(True)
    return [(n, t / reps) for n, t in acc]


def dp_check(ctx, world, rank, bn, peer=False, steps=3, Bg=32):
    """N ranks at Bg/N images each (library NCCL inside the step graph [, SyncBN]) against ONE rank at Bg:
    loss trajectory, final parameters, and bit-identity of the replicas.  Same inputs as
    tests/test_gpu_dist.py (seed 1234).  A ReLU / arg-max decision that flips on a 1e-7 difference makes two correct
    fp32 trajectories part ways discretely (seen with other seeds from the third step on), so the comparison
    stays at three steps."""
    import torch
    import torch.distributed as dist
    from cnn_b200 import nets
    from cnn_b200._lib import check
    from cnn_b200.api import Net
    from cnn_b200.dist import shard_range
    from cnn_b200.synth import synth_images, synth_labels
    spec = nets.alexnet_lite(3, batch_norm=bn)
    init = np.fromfile(os.path.join(ROOT, "tests", "golden", "alexnet_init.model"), np.float32)
    if bn:
        init = nets.insert_bn_params(spec, init)
        check(ctx.L.cnn_dist_set_sync_bn(ctx._h, 1), "cnn_dist_set_sync_bn")
    first, count = shard_range(Bg, world, rank)
    net = Net(ctx, spec, count)
    net.set_params(init)
    if peer:
        net.enable_peer_exchange()
    x = ctx.to_device(synth_images(count, seed=1234, first_image=first))
    lab = ctx.to_device(synth_labels(count, 3, first_image=first), torch.int32)
    losses = []
    for _ in range(steps):
        net.train_step(x, lab, 1e-3, grad_scale=1.0 / Bg, do_update=3)
        ctx.sync()
        losses.append(float(net.loss_from_slab(Bg)))
    params = net.get_params()
    net.close()
    t = torch.from_numpy(params).to(ctx.device)
    mx, mn = t.clone(), t.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    same = bool(torch.equal(mx, mn))
    if bn:
        check(ctx.L.cnn_dist_set_sync_bn(ctx._h, 0), "cnn_dist_set_sync_bn")
    res = None
    if rank == 0:   # the same global batch on one GPU, no collective in the step
        ref = Net(ctx, spec, Bg)
        ref.set_params(init)
        xr = ctx.to_device(synth_images(Bg, seed=1234))
        lr_ = ctx.to_device(synth_labels(Bg, 3), torch.int32)
        rl = []
        for _ in range(steps):
            ref.train_step(xr, lr_, 1e-3, do_update=1)
            ctx.sync()
            rl.append(float(ref.loss_from_slab()))
        rp = ref.get_params()
        ref.close()
        res = {"global_batch": Bg, "steps": steps,
               "loss_rel": float(max(abs(a - b) / max(1.0, abs(b)) for a, b in zip(losses, rl))),
               "params_rel": float(np.abs(params - rp).max() / np.abs(rp).max()),
               "replicas_bit_identical": same, "tolerance": 1e-4}
    dist.barrier()
    return res


def run_ours(a):
    import torch
    import torch.distributed as dist
    from cnn_b200 import api, nets
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
    spec = {"alexnet_lite": lambda: nets.alexnet_lite(3, batch_norm=a.bn), "vgg_style": lambda: nets.vgg_style(3),
            "resnet18_shaped": lambda: nets.resnet18_shaped(3)}[a.net]()
    ctx = Context(local)
    if world > 1 and not a.no_numa:   # pinned staging buffers of this rank on the GPU's NUMA node
        ctx.bind_numa()
    if a.conv_algo != "auto":
        ctx.set_conv_algo({"simt": api.CONV_SIMT, "tcgen05": api.CONV_TCGEN05}[a.conv_algo])
    dtype = "f32"
    if a.precision != "fp32":
        ctx.set_tc_precision({"bf16x3": api.TC_BF16X3, "bf16": api.TC_BF16X1}[a.precision])
        dtype = {"bf16x3": "f32 (bf16x3 split MMA)", "bf16": "bf16 (fp32 accumulate)"}[a.precision]
    net = Net(ctx, spec, B)
    if a.net == "alexnet_lite":
        init = np.fromfile(os.path.join(ROOT, "tests", "golden", "alexnet_init.model"), np.float32)
        net.set_params(nets.insert_bn_params(spec, init) if a.bn else init)
    else:
        net.set_params(nets.scaled_init(spec, seed=0))
    if a.materialize:
        net.set_lazy(False)
    # two resident input batches (> L2 each: 154 MB at B=256) alternate between steps
    xs = [ctx.to_device(synth_images(B, seed=1234 + i, first_image=rank * B)) for i in range(2)]
    lab = ctx.to_device(synth_labels(B, 3, first_image=rank * B), torch.int32)
    from cnn_b200.dist import init_native_dist
    peer = False
    if world > 1:   # the library's own communicator: the gradient exchange sits inside the step's CUDA graph
        init_native_dist(ctx)
        if not a.no_peer:   # one-shot exchange + SGD over NVLink peer memory instead of ncclAllReduce + sgd_kernel
            peer = net.enable_peer_exchange()
        if a.bn:
            from cnn_b200._lib import check
            check(ctx.L.cnn_dist_set_sync_bn(ctx._h, 1), "cnn_dist_set_sync_bn")
    scale = 1.0 / (B * world)
    lr = 1e-3
    upd = 3 if world > 1 else 1

    def step(i):   # one graph launch: fwd + xent + bwd [+ ncclAllReduce of the slab] + SGD
        net.train_step(xs[i & 1], lab, lr, grad_scale=scale, do_update=upd)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n_steps):
        barrier()
        l0 = ctx.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ctx.stream)
        for i in range(n_steps):
            step(i)
        e1.record(ctx.stream)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=ctx.device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / n_steps, ctx.launches - l0

    warm = max(a.warmup, 3)
    for i in range(warm):
        step(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_step, launches = timed(a.steps)
    loss = float(net.loss_from_slab(B * world))
    value = B * world / (ms_step * 1e-3)
    # K steps of < 1 ms are over before nvidia-smi (100 ms period) has sampled twice: the same steps keep
    # running, untimed, for another ~0.4 s so that the clock / throttle record describes this load
    n_obs = max(20, min(2000, int(400.0 / max(ms_step, 1e-3))))
    for i in range(n_obs):
        step(a.steps + i)
    barrier()
    clocks = None
    if rank == 0:
        clocks = sampler.stop()
        clocks["window"] = f"timed region + {n_obs} more identical steps (untimed)"

    # the same step with every reference-visible buffer written (conv/ReLU/pool outputs, int32 mask, image gradient)
    materialized = None
    if not a.materialize and a.net == "alexnet_lite" and not a.bn:
        net.set_lazy(False)
        for i in range(3):
            step(i)
        m_ms, _ = timed(a.steps)
        net.set_lazy(True)
        for i in range(2):
            step(i)
        materialized = {"value": round(B * world / (m_ms * 1e-3), 1), "ms_per_step": round(m_ms, 4),
                        "what": "cnn_net_set_lazy(0): every Layer::get_output buffer, the pool mask and the image gradient "
                                "are written each step, as the reference does"}

    # ---- end to end through the host-buffer C-ABI calls (pinned host memory), every rank its shard ----------
    C_, H_, W_ = 3, 224, 224
    h8 = [torch.from_numpy(np.random.default_rng(5 + i + 100 * rank).integers(0, 256, (B, H_, W_, C_), dtype=np.uint8)).pin_memory()
          for i in range(2)]
    hx = [torch.from_numpy(synth_images(B, seed=4321 + i, first_image=rank * B)).pin_memory() for i in range(2)]
    hl = torch.from_numpy(synth_labels(B, 3, first_image=rank * B)).pin_memory()
    hp = torch.empty(B, net.classes).pin_memory()
    e2e_steps = max(3, min(a.steps, 20))

    def timed_host_loop(run):
        barrier()
        t0 = time.perf_counter()
        run()
        torch.cuda.synchronize()
        dt = torch.tensor([(time.perf_counter() - t0) * 1e3], device=ctx.device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return float(dt.item()) / e2e_steps

    def piped(src):
        # every step: H2D of its batch (copy stream, overlapped with the previous step), [u8: planar /255 on the
        # device,] the step, D2H of loss + probabilities which the host reads in wait_host
        def run():
            net.submit_host(src[0], hl, lr)
            for i in range(1, e2e_steps):
                net.submit_host(src[i & 1], hl, lr)
                net.wait_host(hp)
            net.wait_host(hp)
        return run

    def blocking():
        for i in range(e2e_steps):
            net.train_step_host(hx[i & 1], hl, lr, hp)

    for fn in (piped(h8), piped(hx)) + ((blocking,) if world == 1 else ()):   # warm-up: staging buffers, graphs
        fn()
    u8_ms = timed_host_loop(piped(h8))
    f32_ms = timed_host_loop(piped(hx))
    img_bytes = B * C_ * H_ * W_
    d2h = 4 + B * net.classes * 4
    e2e = {"value": round(B * world / (u8_ms * 1e-3), 1), "unit": UNIT, "h2d_bytes_per_step": img_bytes + B * 4,
           "d2h_bytes_per_step": d2h, "ms_per_step": round(u8_ms, 4), "steps": e2e_steps,
           "call": "cnn_net_train_step_host_submit_u8/_wait per rank: the loader's interleaved u8 HWC image bytes (what "
                   "Tensor3D::read_from_opencv_mat, data_format.cpp:13-23, is handed) from pinned host memory, planar "
                   "x*1.f/255 on the device, depth-2 pipeline, loss + probabilities read back every step"
                   + (", gradient all-reduce inside the step graph" if world > 1 else ""),
           "fp32_images": {"value": round(B * world / (f32_ms * 1e-3), 1), "ms_per_step": round(f32_ms, 4),
                           "h2d_bytes_per_step": img_bytes * 4 + B * 4,
                           "call": "cnn_net_train_step_host_submit/_wait: fp32 CHW host tensors (PCIe-bound: 4x the bytes)"}}
    if world == 1:
        b_ms = timed_host_loop(blocking)
        e2e["blocking_fp32_call"] = {"value": round(B / (b_ms * 1e-3), 1), "ms_per_step": round(b_ms, 4),
                                     "call": "cnn_net_train_step_host"}

    # ---- per-kernel breakdown of the step (CUDA events around every launch of an eager step) ---------------
    roof = pair = breakdown = layer_rows = None
    if not a.no_breakdown:
        rows = profile_step(ctx, net, xs[0], lab, lr, scale, upd)
        if rank == 0:
            hbm, tc, which = peaks()
            costs = layer_costs(spec, B)
            groups = {}
            for name, us in rows:
                tag, _, kern = name.partition(":") if ":" in name else ("", "", name)
                key = (tag, kern)
                g = groups.setdefault(key, [0, 0.0])
                g[0] += 1
                g[1] += us
            tot = sum(v[1] for v in groups.values())
            breakdown = {f"{tag + ':' if tag else ''}{kern}": {"launches": c, "us": round(us, 1), "share": round(us / tot, 4)}
                         for (tag, kern), (c, us) in groups.items()}
            breakdown["_total_us_eager_with_event_gaps"] = round(tot, 1)
            # per (layer, role): time of its kernels against the algorithmic cost of that pass.  Roles of a
            # conv layer's backward: weight gradient (incl. delta packing / partial reduction) and input gradient.
            def role(tag, kern):
                ps = tag[-1]
                if ps == "f":
                    return "forward"
                if "wgrad" in kern or "pack" in kern:      # delta packing feeds both gradients; booked with the weight gradient
                    return "weight gradient"
                if "dgrad" in kern or kern.startswith(("s2_gemm_kernel<true", "s1_gemm_kernel<true", "gather_rows_ws<true")) \
                        or (kern.startswith("gather_gemm_ws<") and kern.split(",")[1].strip() == "true"):
                    return "input gradient"
                return "backward"
            lazy_head = any(k.startswith("head_fwd_kernel") for (_, k) in groups)
            per = {}
            for (tag, kern), (c, us) in groups.items():
                if not (tag.startswith("L") and tag[-1] in "fb") or kern.startswith(("peer_allreduce", "nccl")) or us <= 0:
                    continue          # (the gradient exchange overlaps a layer's backward and carries its tag)
                li, ps = int(tag[1:-1]), tag[-1]
                if lazy_head and (li, ps) == (2, "b"):
                    li = 0          # the lazy head's backward kernel (tagged at the pool layer) is conv1's weight gradient
                e = per.setdefault((li, ps, role(tag, kern)), [0.0, [], 0.0, ""])
                e[0] += us
                e[1].append(kern)
                if us / c > e[2]:
                    e[2], e[3] = us / c, kern       # the pass's dominant kernel and its average launch duration
            best = None
            layer_rows = []
            for (li, ps, rl), (pass_us, kerns, us, top) in per.items():
                fl, by, kind = costs.get((li, ps), (0.0, 0.0, "?"))
                if kind in ("conv", "linear") and ps == "b":
                    if rl == "backward":
                        pass                  # one kernel does both gradients (Linear): full backward cost
                    else:
                        fl, by = fl / 2, by / 2   # one of the two gradient passes
                if by <= 0 or us <= 0 or pass_us <= 0:
                    continue
                ai = fl / by
                tens = kind in ("conv", "linear") and ai > (tc / 3.0) * 1e12 / (hbm * 1e9)
                frac = (fl / (us * 1e-6) / 1e12) / (tc / 3.0) if tens else (by / (us * 1e-6) / 1e9) / hbm
                frac_pass = (fl / (pass_us * 1e-6) / 1e12) / (tc / 3.0) if tens else (by / (pass_us * 1e-6) / 1e9) / hbm
                cand = {"kernel": f"layer {li} {kind} {rl}: {top}",
                        "pass_kernels": "+".join(sorted(set(kerns))), "pass_us": round(pass_us, 1), "frac_pass": round(frac_pass, 4),
                        "bound": "tensor" if tens else "hbm",
                        "achieved": round(fl / (us * 1e-6) / 1e12, 2) if tens else round(by / (us * 1e-6) / 1e9, 1),
                        "peak": round(tc / 3.0, 1) if tens else hbm, "unit": "TFLOP/s" if tens else "GB/s",
                        "frac": round(frac, 4), "traffic": None,
                        "peak_source": which + (", bf16 sustained / 3 (3-pass split MMA)" if tens else ""),
                        "algorithmic_bytes": by, "flops": fl, "launch_us": round(us, 1),
                        "arithmetic_intensity": round(ai, 1)}
                # per (layer, pass): achieved / min(tensor ceiling, AI x HBM peak), SURVEY 8(d)'s formula, over ALL kernels of
                # the pass (packing, partial reductions included)
                ceil_tf = min(tc / 3.0, ai * hbm * 1e9 / 1e12) if fl > 0 else None
                layer_rows.append({"layer": li, "kind": kind, "pass": rl, "us": round(pass_us, 1),
                                   "tflops": round(fl / (pass_us * 1e-6) / 1e12, 2) if fl > 0 else None,
                                   "gbs": round(by / (pass_us * 1e-6) / 1e9, 1),
                                   "bound": "tensor" if tens else "hbm",
                                   "frac_of_roofline": round((fl / (pass_us * 1e-6) / 1e12) / ceil_tf, 4) if fl > 0 and tens
                                   else round((by / (pass_us * 1e-6) / 1e9) / hbm, 4)})
                if lazy_head and li == 0 and kind == "conv":
                    cand["note"] = ("algorithmic bytes are SURVEY 8(d)'s figure for this conv pass (X + Y resp. X + delta in fp32); the "
                                    "fused lazy-head kernel also does the ReLU + max-pool work and moves fewer bytes than that "
                                    "(see traffic)")
                if best is None or us > best[0]:
                    best = (us, cand)
            if best:
                roof = best[1]
                try:   # ncu dram bytes of the same kernels, captured in the same gpurun as the final bench (tools/gpu_final.sh)
                    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                        tr = json.load(f)
                    names = [roof["kernel"].split(": ", 1)[1].split("<")[0]]
                    hit = [v for k, v in tr.items() if k.split("<")[0] in names]
                    roof["traffic"] = int(sum(hit)) if hit else None
                except Exception:
                    pass
            if lazy_head:
                t_f = sum(us for (tag, k), (c, us) in groups.items() if k.startswith("head_fwd_kernel"))
                t_w = sum(us for (tag, k), (c, us) in groups.items() if k.startswith("head_wgrad_kernel") or k.startswith("thin_wgrad_reduce"))
                by = 2 * costs[(0, "f")][1]
                fl = 2 * costs[(0, "f")][0]
                pair = {"what": "north-star pair, conv1 3->16 forward + weight gradient at this batch (SURVEY 8d: X+Y and X+delta "
                                "compulsory bytes); the two kernels also do the ReLU + max-pool forward / backward",
                        "algorithmic_bytes": by, "flops": fl, "us": round(t_f + t_w, 1),
                        "achieved_gbs": round(by / ((t_f + t_w) * 1e-6) / 1e9, 1), "frac_of_hbm_peak": round(by / ((t_f + t_w) * 1e-6) / 1e9 / hbm, 4),
                        "tflops": round(fl / ((t_f + t_w) * 1e-6) / 1e12, 2)}

    check_res = None
    if world > 1 and not a.no_dp_check:
        check_res = {"plain": dp_check(ctx, world, rank, False, peer), "sync_bn": dp_check(ctx, world, rank, True, peer)}
        if rank == 0:
            check_res = {**check_res["plain"], "sync_bn": check_res["sync_bn"]}

    line = None
    if rank == 0:
        cpu = None
        if world == 1 and not a.no_cpu:
            kind, rate, dt = cpu_reference_rate(1, 4, a.cpu_steps, 1)
            cpu = {"value": round(rate, 3), "unit": UNIT, "cores": 1, "kind": kind,
                   "sample": f"{a.cpu_steps} train steps at batch 4 (the reference's own batch, cnn.cpp:36), "
                             f"single thread, {dt:.1f}s; host has {os.cpu_count()} cores"}
        lazy_on = not a.materialize and a.net == "alexnet_lite" and not a.bn
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": warm, "ms_per_step": round(ms_step, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": {"workload": f"{a.net}{' +BatchNorm' if a.bn else ''} train step (fwd+xent+bwd+SGD), "
                                   f"batch {B}/GPU x {world} GPU, 3x224x224 fp32, "
                                   f"{'reference-seed' if a.net == 'alexnet_lite' else 'fan-in scaled random'} init",
                       "global_batch": B * world, "parallelism": f"dp{world}", "conv_algo": a.conv_algo,
                       "allreduce": "none" if world == 1 else
                                    ("one-shot peer-memory exchange fused with SGD (every rank reads all slabs over NVLink, rank-order sum), "
                                     "inside the step graph" if peer else
                                     "one ncclAllReduce of the gradient slab issued by the library inside the step graph"),
                       "lazy_head": ("on (default API): conv1/ReLU/pool outputs, pool mask and image gradient are re-created on "
                                     "demand, see `materialized` for the step that writes them all") if lazy_on else "off",
                       "l2": f"two alternating resident input batches of {img_bytes * 4 / 1e6:.0f} MB each; per-step activations exceed the 126 MB L2",
                       "cuda_graph": True},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "loss_after": loss,
            "materialized": materialized, "roofline": roof, "north_star_pair": pair, "cpu_baseline": cpu,
            "dp_check": check_res, "layers": sorted(layer_rows, key=lambda r: (r["layer"], r["pass"])) if layer_rows else None,
            "breakdown": breakdown,
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
    ap.add_argument("--net", default="alexnet_lite", choices=["alexnet_lite", "vgg_style", "resnet18_shaped"])
    ap.add_argument("--bn", action="store_true", help="alexnet_lite with BatchNorm2D after every conv (SyncBN when N>1)")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16x3", "bf16"],
                    help="tensor-core operand mode of the generic conv kernels (fp32 = TF32x3 split, reference parity)")
    ap.add_argument("--materialize", action="store_true", help="cnn_net_set_lazy(0) for the headline value")
    ap.add_argument("--no-dp-check", action="store_true")
    ap.add_argument("--no-numa", action="store_true", help="N>1: do not bind each rank to its GPU's NUMA node")
    ap.add_argument("--no-peer", action="store_true", help="N>1: keep ncclAllReduce + SGD kernel instead of the peer-memory exchange")
    ap.add_argument("--conv-algo", default="auto", choices=["auto", "simt", "tcgen05"])
    ap.add_argument("--cpu-steps", type=int, default=40)
    ap.add_argument("--ref-procs", type=int, default=0, help="reference arm processes (0 = all cores)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-breakdown", action="store_true")
    a = ap.parse_args()
    with OneLineStdout() as OUT:
        try:
            if a.impl == "reference":
                run_reference(a)
            else:
                run_ours(a)
        except BaseException:
            # one rank failing must end the job, not leave the others waiting in a collective until a time limit
            import traceback
            traceback.print_exc()
            sys.stderr.flush()
            if int(os.environ.get("WORLD_SIZE", "1")) > 1:
                os._exit(1)
            raise


if __name__ == "__main__":
    main()
