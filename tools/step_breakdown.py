#!/usr/bin/env python
"""Per-kernel-family time breakdown of one training step (CUDA events around every C-ABI call of an eager run).
Usage (GPU box):  python tools/step_breakdown.py [--net r50ppm|hrnet|r101upernet] [--batch 2] [--crop 512] [--top 25]
Writes a table to stdout; used to decide what to optimise next and to cross-check ncu launch lists."""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "semantic-segmentation-pytorch_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--crop", type=int, default=512)
    ap.add_argument("--top", type=int, default=30)
    ap.add_argument("--detail", action="store_true", help="list every conv GEMM launch with shape and TFLOP/s")
    ap.add_argument("--replay-only", action="store_true", help="only print the CUDA-graph replay time")
    ap.add_argument("--net", default="r50ppm", choices=["r50ppm", "hrnet", "r101upernet"],
                    help="r50ppm = the bench network (config 3); hrnet = HRNetV2+C1 (config 5); r101upernet = config 4")
    args = ap.parse_args()
    from mit_semseg.engine import ops
    from mit_semseg.engine.program import SegProgram
    dev = torch.device("cuda", 0)
    stride = 8
    if args.net == "r50ppm":
        seg = bench.build_model(dev)
    else:
        import torch.nn as nn
        from mit_semseg.models import ModelBuilder, SegmentationModule
        from mit_semseg.models import hrnet as HR, models as M, resnet as R
        torch.manual_seed(304)
        if args.net == "hrnet":
            enc, dec = HR.hrnetv2(pretrained=False), ModelBuilder.build_decoder("c1", fc_dim=720, num_class=150)
        else:
            enc = M.Resnet(R.resnet101(pretrained=False))
            dec = ModelBuilder.build_decoder("upernet", fc_dim=2048, num_class=150)
        stride = 4
        seg = SegmentationModule(enc, dec, nn.NLLLoss(ignore_index=-1), None).to(dev).train()
    feed = bench.synth_batch(args.batch, args.crop, args.crop, stride, 304)
    prog = SegProgram(seg, tuple(feed["img_data"].shape), training=True, with_grad=True)
    prog.load_inputs(feed["img_data"].to(dev), feed["seg_label"].to(dev))
    prog.run_eager()
    prog.run_eager()
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream()
    if args.replay_only:
        prog.capture()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(3):
            t0.record(stream)
            for _ in range(10):
                prog.run()
            t1.record(stream)
            torch.cuda.synchronize()
            best = min(best, t0.elapsed_time(t1) / 10)
        print("CUDA-graph replay: %.3f ms / step (best of 3x10)" % best)
        return
    recs = []
    names = [n for n in dir(ops) if callable(getattr(ops, n)) and not n.startswith("_") and
             n not in ("act", "make_geom", "conv_taps", "conv_s2_taps", "parity_planes")]
    orig = {n: getattr(ops, n) for n in names}

    def wrap(name, fn):
        def w(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            r = fn(*a, **k)
            e1.record(stream)
            info = ""
            if name in ("conv_igemm", "conv_wgrad"):
                g = a[0]
                k_tot = sum((g.srcs[0].c if g.tap_src[t] >= 0 else sum(g.srcs[i].c for i in range(g.nsrc)))
                            for t in range(g.ntaps))
                o = a[3] if name == "conv_igemm" else a[1]
                cout = a[2]
                fl = 2.0 * o.shape[0] * o.shape[1] * o.shape[2] * cout * k_tot
                info = (fl, "%dx%dx%d cout=%d K=%d taps=%d" % (o.shape[0], o.shape[1], o.shape[2], cout, k_tot, g.ntaps))
            recs.append((name, e0, e1, info))
            return r
        return w

    for n in names:
        setattr(ops, n, wrap(n, orig[n]))
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    prog.run_eager()
    t1.record(stream)
    torch.cuda.synchronize()
    for n in names:
        setattr(ops, n, orig[n])
    agg = collections.OrderedDict()
    for name, e0, e1, info in recs:
        d = agg.setdefault(name, [0, 0.0, 0.0])
        d[0] += 1
        d[1] += e0.elapsed_time(e1)
        if info:
            d[2] += info[0]
    total = sum(v[1] for v in agg.values())
    print("eager step %.3f ms; sum of bracketed kernels %.3f ms; %d launches" % (t0.elapsed_time(t1), total, len(recs)))
    print("%-22s %6s %10s %7s %10s" % ("op", "calls", "ms", "%", "TFLOP/s"))
    for name, (cnt, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:args.top]:
        print("%-22s %6d %10.3f %6.1f%% %10s" % (name, cnt, ms, 100 * ms / total, ("%.1f" % (fl / ms / 1e9)) if fl else ""))
    if args.detail:
        print("\nper-launch GEMMs (slowest first):")
        gem = [(e0.elapsed_time(e1), name, info) for name, e0, e1, info in recs if info]
        for ms, name, (fl, desc) in sorted(gem, key=lambda x: -x[0])[:60]:
            print("%-10s %8.3f ms %8.1f TFLOP/s  %s" % (name, ms, fl / ms / 1e9, desc))
    # true per-kernel GPU durations (CUPTI via torch.profiler; unaffected by CPU launch gaps)
    try:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            prog.run_eager()
            torch.cuda.synchronize()
        kagg = collections.OrderedDict()
        for ev in prof.events():
            if ev.device_type.name == "CUDA" or "cuda" in str(ev.device_type).lower():
                d = kagg.setdefault(ev.name[:70], [0, 0.0])
                d[0] += 1
                d[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
        ktot = sum(v[1] for v in kagg.values())
        print("\nprofiler: %d kernel names, total GPU busy %.3f ms" % (len(kagg), ktot / 1e3))
        for name, (cnt, us) in sorted(kagg.items(), key=lambda kv: -kv[1][1])[:args.top]:
            print("%-70s %5d %9.3f ms %5.1f%%" % (name, cnt, us / 1e3, 100 * us / ktot))
    except Exception as exc:  # profiler availability varies
        print("profiler unavailable:", exc)
    # graph replay for comparison
    prog.capture()
    torch.cuda.synchronize()
    t0.record(stream)
    for _ in range(10):
        prog.run()
    t1.record(stream)
    torch.cuda.synchronize()
    print("CUDA-graph replay: %.3f ms / step" % (t0.elapsed_time(t1) / 10))


if __name__ == "__main__":
    main()
