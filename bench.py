#!/usr/bin/env python
"""bench.py — training images/sec of ResNet50dilated + PPM_deepsup on synthetic 3x512x512 batches (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run, one rank/GPU)
    python bench.py --impl reference --gpus N --steps K ...   (the reference's CPU path = the oracle port, host cores)

One "step" = forward + backward (+ SyncBN / gradient-bucket NCCL all-reduces when N > 1) + SGD update of all
51.6 M parameters on a per-GPU batch of 2 x 3 x 512 x 512 (labels 2 x 64 x 64): config 3 of BASELINE.json
("ResNet50dilated + PPM_deepsup training bf16, SyncBN allreduce, crop 512, 2 imgs/GPU"), weak scaling.

Prints ONE JSON line (rank 0).  `value`: images/s with inputs resident in HBM, the step replayed as one CUDA graph,
device-timed with CUDA events (max over ranks).  `e2e`: the same metric through the public API a user calls
(SegmentationModule(feed_dict) -> loss.backward() -> optimizer.step(), train.py:41-48) with pinned HOST inputs copied
every step and the loss read back every step.  `roofline`: the tcgen05 conv kernels' achieved FLOP/s over their own
CUDA-event-timed launches (an eager pass right after the timed region, same buffers) against the measured bf16 peak.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "semantic-segmentation-pytorch_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

ENC_ARCH, DEC_ARCH, FC_DIM, NUM_CLASS = "resnet50dilated", "ppm_deepsup", 2048, 150
BATCH, CROP, LABEL_STRIDE = 2, 512, 8
TRAIN_GFLOP_PER_IMG = 1224.2  # fwd + dgrad + wgrad conv FLOPs per image
LR, MOMENTUM, WD = 0.02, 0.9, 1e-4  # config/ade20k-resnet50dilated-ppm_deepsup.yaml:17-27


def synth_batch(n, h, w, label_stride, seed, num_class=NUM_CLASS):
    """Synthetic ADE20K-shaped batch: img ~ N(0,1), labels uniform in {-1 .. num_class-1} (-1 = ignore).
    Same generator as oracle.segnet_oracle.synth_batch, restated here so the GPU arm never imports the oracle."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    img = torch.randn(n, 3, h, w, generator=g)
    label = torch.randint(-1, num_class, (n, h // label_stride, w // label_stride), generator=g)
    return {"img_data": img, "seg_label": label}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1381.2)), d.get("bf16_tflops", 1665.0), "measured"
    return 1400.0, 1590.0, "fallback"


def group_weight(module):
    """train.py:92-112: weight decay on conv/linear weights only."""
    decay, no_decay = [], []
    for m in module.modules():
        if isinstance(m, (nn.Linear, nn.modules.conv._ConvNd)):
            decay.append(m.weight)
            if m.bias is not None:
                no_decay.append(m.bias)
        elif isinstance(m, nn.modules.batchnorm._BatchNorm):
            if m.weight is not None:
                no_decay.append(m.weight)
            if m.bias is not None:
                no_decay.append(m.bias)
    return [dict(params=decay), dict(params=no_decay, weight_decay=0.0)]


def build_model(device):
    from mit_semseg.models import ModelBuilder, SegmentationModule
    from mit_semseg.models import models as M, resnet as R
    torch.manual_seed(304)  # config/defaults.py:75
    enc = M.ResnetDilated(R.resnet50(pretrained=False), 8)  # reference initialisers, random weights (no network)
    dec = ModelBuilder.build_decoder(DEC_ARCH, fc_dim=FC_DIM, num_class=NUM_CLASS)
    seg = SegmentationModule(enc, dec, nn.NLLLoss(ignore_index=-1), 0.4)
    return seg.to(device).train()


def make_optimizers(seg, fused=False):
    """train.py:115-127: one SGD per net. fused=True: the engine's one-launch multi-tensor SGD (same arithmetic)."""
    if fused:
        from mit_semseg.engine.optim import FusedSGD
        return [FusedSGD(group_weight(net), lr=LR, momentum=MOMENTUM, weight_decay=WD) for net in (seg.encoder, seg.decoder)]
    return [torch.optim.SGD(group_weight(net), lr=LR, momentum=MOMENTUM, weight_decay=WD)
            for net in (seg.encoder, seg.decoder)]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc = gpu_index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], None, set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    return rank, local, world


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(value, world):
    if world == 1:
        return value
    import torch.distributed as dist
    t = torch.tensor([value], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    rank, local, world = dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    from mit_semseg.engine import _C, ops
    from mit_semseg.engine.program import SegProgram

    seg = build_model(dev)
    opts = make_optimizers(seg, fused=True)   # device-resident arm: engine optimizer (2 launches / step)
    feed = synth_batch(BATCH, CROP, CROP, LABEL_STRIDE, 304 + rank, NUM_CLASS)
    img_h, lab_h = feed["img_data"].pin_memory(), feed["seg_label"].pin_memory()
    img_d, lab_d = img_h.to(dev), lab_h.to(dev)

    # ---------------- device-resident arm: one CUDA graph = weight prep + fwd + bwd + all-reduces + grad re-layout
    prog = SegProgram(seg, tuple(img_d.shape), training=True, with_grad=True)
    prog.load_inputs(img_d, lab_d)
    launches_per_step = prog.num_launches()
    prog.capture()
    grads = prog.param_grads()
    for p in seg.parameters():
        p.grad = grads[p]

    def step():
        prog.run()
        for o in opts:
            o.step()

    for _ in range(args.warmup):
        step()
    barrier(world)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier(world)
    clocks = sampler.stop() if rank == 0 else None
    ms = max_over_ranks(ev0.elapsed_time(ev1), world)
    ms_per_step = ms / args.steps
    value = BATCH * world * args.steps / (ms / 1e3)
    loss_dev = prog.out[0].item()

    # ---------------- roofline of the tensor-core kernels: CUDA events around every conv GEMM launch (eager pass)
    # every rank runs the pass (the program contains the SyncBN / bucket all-reduces); rank 0 reports its own numbers
    roof = conv_roofline(prog, world)
    barrier(world)

    # ---------------- end-to-end arm: the public API with host inputs (H2D) and a loss read-back (D2H) every step.
    # Twice: with the optimizer this package ships (mit_semseg.engine.optim.FusedSGD, a torch.optim.Optimizer with
    # torch.optim.SGD's arithmetic: the headline `e2e`) and with exactly train.py's torch.optim.SGD pair (`e2e.torch_optim_sgd`).
    feed_host = {"img_data": img_h, "seg_label": lab_h}

    def e2e_run(fused):
        for p in seg.parameters():
            p.grad = None
        e_opts = make_optimizers(seg, fused=fused)

        def e2e_step():
            seg.zero_grad()
            # pinned HOST tensors go straight into the public call: SegmentationModule copies them (cudaMemcpyAsync) into the
            # step's static device buffers - this is the H2D traffic counted below
            loss, acc = seg(feed_host)
            loss = loss.mean()
            loss.backward()
            for o in e_opts:
                o.step()
            return loss.item()

        for _ in range(max(3, args.warmup)):
            e2e_step()
        barrier(world)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            last_ = e2e_step()
        e1.record()
        barrier(world)
        ms_ = max_over_ranks(e0.elapsed_time(e1), world)
        return ms_, BATCH * world * args.steps / (ms_ / 1e3), last_

    e2e_ms, e2e_value, last = e2e_run(fused=True)
    e2e_ms_t, e2e_value_t, last_t = e2e_run(fused=False)

    prog.graph = None
    seg.__dict__.pop("_b200_programs", None)
    if rank != 0:
        return
    sustained, burst, which = measured_peaks()
    out = {
        "metric": "ResNet50dilated+PPM_deepsup training images/sec (synthetic 3x512x512)",
        "value": round(value, 3), "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "configs[2]: ResNet50dilated+PPM_deepsup train step (fwd+bwd+SGD), %d x 3x%dx%d per GPU, "
                               "labels %dx%d, 150 classes" % (BATCH, CROP, CROP, CROP // 8, CROP // 8),
                   "global_batch": BATCH * world, "parallelism": "dp%d" % world,
                   "l2": "per-step working set (~3 GB activations + 0.6 GB weights/grads) exceeds the 126 MB L2",
                   "weights": "reference initialisers, seed 304, random (no checkpoints offline)",
                   "loss_last": round(loss_dev, 5)},
        "clocks": clocks,
        "e2e": {"value": round(e2e_value, 3), "unit": "images/s", "ms_per_step": round(e2e_ms / args.steps, 4),
                "h2d_bytes_per_step": img_h.numel() * 4 + lab_h.numel() * 8, "d2h_bytes_per_step": 4,
                "loss_last": round(last, 5),
                "api": "SegmentationModule(host feed) -> loss.mean().backward() -> FusedSGD.step() x2 -> loss.item()",
                "torch_optim_sgd": {"value": round(e2e_value_t, 3), "ms_per_step": round(e2e_ms_t / args.steps, 4),
                                    "loss_last": round(last_t, 5)}},
        "gpu_launches": (launches_per_step + 2) * args.steps,
        "launches_per_step": launches_per_step + 2,
        "model_flops_frac": round(value / world * TRAIN_GFLOP_PER_IMG / 1e3 / sustained, 4),
        "roofline": roof,
    }
    if world == 1 and not args.no_gpu_context:
        out["gpu_context"] = gpu_context(dev)
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(max_seconds=25.0)
    print(json.dumps(out))


def gpu_context(dev, steps=10):
    """Same-box context number ("the bar on the same box is PyTorch eager + cuDNN"): the reference's arithmetic
    (the oracle port: same torch ops in the reference's order) executed by stock PyTorch ON THE GPU - fp32 with TF32 off, and
    bf16 autocast with channels_last inputs - for the same 2x3x512x512 training step (fwd + bwd + torch.optim.SGD).
    Bench-only code: nothing of the product path is involved and nothing here is imported by the package."""
    from oracle import segnet_oracle as O
    res = {"what": "oracle port (reference ops, stock PyTorch eager + cuDNN) on the same GPU, %d x 3x%dx%d train step"
                   % (BATCH, CROP, CROP)}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    try:
        torch.backends.cudnn.benchmark = True
        feed = {k: v.to(dev) for k, v in O.synth_batch(BATCH, CROP, CROP, LABEL_STRIDE, 304, NUM_CLASS).items()}
        for tag, tf32, autocast in (("fp32_tf32_off", False, False), ("bf16_autocast_channels_last", True, True)):
            torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = tf32
            esd = O.synth_state_dict(O.encoder_param_shapes(ENC_ARCH), 304)
            dsd = O.synth_state_dict(O.decoder_param_shapes(DEC_ARCH, FC_DIM), 305)
            e = {k: v.to(dev).requires_grad_(v.is_floating_point() and "running" not in k) for k, v in esd.items()}
            d = {k: v.to(dev).requires_grad_(v.is_floating_point() and "running" not in k) for k, v in dsd.items()}
            if autocast:
                for sd in (e, d):
                    for k, v in sd.items():
                        if v.dim() == 4:
                            v.data = v.data.contiguous(memory_format=torch.channels_last)
            params = [v for v in list(e.values()) + list(d.values()) if v.requires_grad]
            opt = torch.optim.SGD(params, lr=LR, momentum=MOMENTUM, weight_decay=WD)
            st = O.BNState(training=True)
            f = dict(feed)
            if autocast:
                f["img_data"] = f["img_data"].contiguous(memory_format=torch.channels_last)

            def step():
                opt.zero_grad()
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                    loss, acc = O.segmentation_forward(f, e, d, ENC_ARCH, DEC_ARCH, st, 0.4)
                loss.backward()
                opt.step()
                return loss
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(steps):
                loss = step()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / steps
            res[tag] = {"ms_per_step": round(ms, 3), "images_per_s": round(BATCH / ms * 1e3, 2), "loss_last": round(loss.item(), 5)}
            del e, d, params, opt
            torch.cuda.empty_cache()
    except Exception as exc:   # context only: never fail the bench line over it
        res["error"] = "%s: %s" % (type(exc).__name__, exc)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old
    return res


def conv_roofline(prog, world):
    """Re-run the program eagerly with a CUDA-event pair around every tcgen05 GEMM launch (conv forward, data
    gradient, weight gradient); achieved = their algorithmic FLOPs / their summed durations."""
    from mit_semseg.engine import ops
    sustained, burst, which = measured_peaks()
    records = []
    orig_igemm, orig_wgrad = ops.conv_igemm, ops.conv_wgrad

    def flops_geom(geom, cout, n, h, w):
        k = 0
        for t in range(geom.ntaps):
            k += geom.srcs[0].c if geom.tap_src[t] >= 0 else sum(geom.srcs[i].c for i in range(geom.nsrc))
        return 2.0 * n * h * w * cout * k

    def timed(kind, fn, fl):
        stream = torch.cuda.current_stream()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        records.append((kind, a, b, fl))

    def igemm(geom, w_bf16, cout, out, **kw):
        n, h, w = out.shape[0], out.shape[1], out.shape[2]
        timed("igemm", lambda: orig_igemm(geom, w_bf16, cout, out, **kw), flops_geom(geom, cout, n, h, w))
        return out

    def wgrad(geom, dy, cout, dw):
        n, h, w = dy.shape[0], dy.shape[1], dy.shape[2]
        timed("wgrad", lambda: orig_wgrad(geom, dy, cout, dw), flops_geom(geom, cout, n, h, w))
        return dw

    orig_bnbwd = ops.conv_igemm_bnbwd

    def igemm_bnbwd(geom, w_bf16, cout, out, *a, **kw):
        # data-gradient launch with the producer's BN-backward reduction fused in its epilogue: same GEMM FLOPs
        n, h, w = out.shape[0], out.shape[1], out.shape[2]
        timed("igemm", lambda: orig_bnbwd(geom, w_bf16, cout, out, *a, **kw), flops_geom(geom, cout, n, h, w))
        return out

    orig_bnbwd_res = ops.conv_igemm_bnbwd_res

    def igemm_bnbwd_res(geom, w_bf16, cout, out, *a, **kw):   # the same for residual-block outputs (mask from the saved output)
        n, h, w = out.shape[0], out.shape[1], out.shape[2]
        timed("igemm", lambda: orig_bnbwd_res(geom, w_bf16, cout, out, *a, **kw), flops_geom(geom, cout, n, h, w))
        return out

    ops.conv_igemm, ops.conv_wgrad, ops.conv_igemm_bnbwd, ops.conv_igemm_bnbwd_res = igemm, wgrad, igemm_bnbwd, igemm_bnbwd_res
    prog.serial = True  # weight gradients inline on the main stream: one kernel at a time between each event pair
    try:
        stream = torch.cuda.current_stream()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # keep the GPU busy for ~40 ms first so the whole eager pass is queued ahead of execution: the event pairs then
        # bracket kernel execution only, not CPU launch latency
        torch.cuda._sleep(int(0.04 * 1.9e9))
        t0.record(stream)
        prog.run_eager()
        t1.record(stream)
        torch.cuda.synchronize()
    finally:
        prog.serial = False
        ops.conv_igemm, ops.conv_wgrad, ops.conv_igemm_bnbwd = orig_igemm, orig_wgrad, orig_bnbwd
        ops.conv_igemm_bnbwd_res = orig_bnbwd_res
    tot = {"igemm": [0.0, 0.0, 0], "wgrad": [0.0, 0.0, 0]}
    top = None
    for kind, a, b, fl in records:
        dt = a.elapsed_time(b) * 1e-3
        tot[kind][0] += fl
        tot[kind][1] += dt
        tot[kind][2] += 1
        if kind == "igemm" and (top is None or fl > top[0]):
            top = (fl, dt)
    fl = tot["igemm"][0] + tot["wgrad"][0]
    sec = tot["igemm"][1] + tot["wgrad"][1]
    n_launch = tot["igemm"][2] + tot["wgrad"][2]
    # dominant kernel FAMILY = the two tcgen05 GEMM kernels (conv forward / data gradient, weight gradient): achieved =
    # algorithmic FLOPs of all their launches of one step / the sum of their CUDA-event durations, against the SUSTAINED
    # measured bf16 peak (they run inside a long step). The largest single launch (decoder.conv_last.0 forward, 3x3
    # 4096->512 over the virtual concat) is reported next to it against the burst peak; its DRAM traffic is not measured
    # by this run (an ncu capture of exactly this launch is in profiles/, see `traffic_profile`).
    top_tflops = top[0] / top[1] / 1e12
    is_conv_last = abs(top[0] - 2.0 * 2 * 64 * 64 * 512 * 36864) < 1.0
    fam = fl / sec / 1e12
    return {"bound": "tensor", "kernel": "igemm_kernel + wgrad_kernel (tcgen05 implicit-GEMM conv fwd / dgrad / wgrad), "
                                         "all %d launches of one step" % n_launch,
            "achieved": round(fam, 2), "peak": sustained, "peak_kind": "%s bf16_tflops_sustained" % which,
            "unit": "TFLOP/s", "frac": round(fam / sustained, 4), "traffic": None,
            "flops_per_step": fl, "gemm_ms_per_step": round(sec * 1e3, 3),
            "igemm_tflops": round(tot["igemm"][0] / max(tot["igemm"][1], 1e-9) / 1e12, 2),
            "wgrad_tflops": round(tot["wgrad"][0] / max(tot["wgrad"][1], 1e-9) / 1e12, 2),
            "largest_launch": {"what": "decoder.conv_last.0 fwd" if is_conv_last else "largest igemm launch",
                               "achieved": round(top_tflops, 2), "peak": burst,
                               "peak_kind": "%s bf16_tflops (burst: one kernel timed alone)" % which,
                               "frac": round(top_tflops / burst, 4), "flops_per_launch": top[0],
                               "launch_ms": round(top[1] * 1e3, 4),
                               "algorithmic_bytes": 113246208 if is_conv_last else None,
                               "traffic_profile": "profiles/r1_summary.md: 143.6 MB dram read+write for this launch (ncu --set full)"
                               if is_conv_last else None},
            "eager_step_ms": round(t0.elapsed_time(t1), 3)}


# ------------------------------------------------------------------------------------------------ CPU arms
def cpu_threads():
    """oneDNN convolutions scale poorly past ~32 threads on many-core hosts (128-thread runs were 3x slower on the
    B200 box); use every core up to 32."""
    return max(1, min(os.cpu_count() or 1, 32))


def oracle_train_setup(n, crop, threads):
    from oracle import segnet_oracle as O
    torch.set_num_threads(threads)
    esd = O.synth_state_dict(O.encoder_param_shapes(ENC_ARCH), 304)
    dsd = O.synth_state_dict(O.decoder_param_shapes(DEC_ARCH, FC_DIM), 305)
    e = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in esd.items()}
    d = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in dsd.items()}
    params = [v for v in list(e.values()) + list(d.values()) if v.requires_grad]
    opt = torch.optim.SGD(params, lr=LR, momentum=MOMENTUM, weight_decay=WD)
    feed = O.synth_batch(n, crop, crop, LABEL_STRIDE, 304, NUM_CLASS)
    st = O.BNState(training=True)

    def step():
        opt.zero_grad()
        loss, acc = O.segmentation_forward(feed, e, d, ENC_ARCH, DEC_ARCH, st, 0.4)
        loss.backward()
        opt.step()
        return loss.item()

    def fwd_only():
        with torch.no_grad():
            O.segmentation_forward(feed, e, d, ENC_ARCH, DEC_ARCH, O.BNState(False), None, segSize=(crop, crop))

    return step, fwd_only


def cpu_baseline(max_seconds=25.0):
    """The reference's CPU path (oracle port: same torch-CPU ops in the reference's order) on this box's host cores:
    bounded sample = one warm-up + up to 3 training steps of the same 2x3x512x512 batch."""
    cores = cpu_threads()
    step, fwd_only = oracle_train_setup(BATCH, CROP, cores)
    t0 = time.perf_counter()
    step()
    warm = time.perf_counter() - t0
    times = []
    while len(times) < 3 and (sum(times) + warm) < max_seconds:
        t = time.perf_counter()
        step()
        times.append(time.perf_counter() - t)
    t = time.perf_counter()
    fwd_only()
    fwd = time.perf_counter() - t
    best = min(times) if times else warm
    return {"value": round(BATCH / best, 4), "unit": "images/s", "cores": cores, "kind": "port",
            "sample": "%d train steps (fwd+bwd+SGD, fp32) of %dx3x%dx%d after 1 warm-up; min step %.2f s; "
                      "eval forward (softmax @%d^2) %.2f s = %.2f img/s" % (len(times), BATCH, CROP, CROP, best, CROP, fwd,
                                                                            BATCH / fwd)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = cpu_threads()
    n, crop = BATCH, CROP
    step, _ = oracle_train_setup(n, crop, cores)
    t0 = time.perf_counter()
    step()
    t_first = time.perf_counter() - t0
    # keep the whole run within a few minutes by timing FEWER steps, never a smaller batch: a 1-image train-mode step does
    # not exist for this network (F.batch_norm on the [1,512,1,1] pyramid branch raises)
    budget = 280.0
    warm = max(0, min(args.warmup - 1, int(0.2 * budget / t_first)))
    steps = max(1, min(args.steps, int((budget - (1 + warm) * t_first) / t_first)))
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        last = step()
    dt = time.perf_counter() - t0
    steps_requested, args.steps = args.steps, steps
    value = n * args.steps / dt
    out = {"impl": "reference",
           "metric": "ResNet50dilated+PPM_deepsup training images/sec (synthetic 3x512x512)",
           "value": round(value, 4), "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": round(dt / args.steps * 1e3, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": "configs[2]: ResNet50dilated+PPM_deepsup train step (fwd+bwd+SGD) on the host CPU, "
                                  "%d x 3x%dx%d per step" % (n, crop, crop), "global_batch": n, "parallelism": "cpu",
                      "steps_requested": steps_requested},
           "cpu_baseline": {"value": round(value, 4), "unit": "images/s", "cores": cores, "kind": "port",
                            "sample": "%d steps of %dx3x%dx%d, %d host threads (of %d logical CPUs); the reference is pure Python over "
                                      "torch CPU ops and cannot travel to the box, so the oracle port (same ops, same "
                                      "order, bit-identical in the build container) stands in" % (args.steps, n, crop,
                                                                                                   crop, cores,
                                                                                                   os.cpu_count() or 1)},
           "e2e": {"value": round(value, 4), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "loss_last": round(last, 5)}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-context", action="store_true", help="skip the same-box PyTorch eager + cuDNN timing")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device - the B200 engine has no CPU path (use --impl reference for the "
                             "CPU arm)")
        run_gpu(args)
        if int(os.environ.get("WORLD_SIZE", "1")) > 1:
            shutdown_distributed()


def shutdown_distributed():
    """Tear the process group down without hanging: CUDA graphs that captured NCCL kernels keep the communicator busy
    at exit (observed: destroy_process_group never returns), so release them first and never wait more than a few
    seconds - the JSON line is already printed."""
    import gc
    import threading
    import torch.distributed as dist
    sys.stdout.flush()
    sys.stderr.flush()
    threading.Timer(15.0, lambda: os._exit(0)).start()
    try:
        torch.cuda.synchronize()
        dist.barrier()
        gc.collect()
        torch.cuda.synchronize()
        dist.destroy_process_group()
    finally:
        os._exit(0)


if __name__ == "__main__":
    main()
