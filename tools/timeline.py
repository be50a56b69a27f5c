#!/usr/bin/env python
"""Device timeline of ONE CUDA-graph replay of the training step (CUPTI kernel records via torch.profiler).

Unlike the ncu launch list (cold caches, serialised) this shows the step as it really runs: warm L2, programmatic
dependent launch, side / branch streams. Per kernel: start, duration, stream, grid; summary: per-family busy time, time
the main stream spends idle between kernels, how much of the step at least one kernel is running.

    python tools/timeline.py [--net r50ppm|hrnet] [--out gpurun_out/timeline.csv] [--eager]
"""
import argparse
import collections
import json
import os
import re
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "semantic-segmentation-pytorch_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("sseg::", "")
    return name[:60]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--net", default="r50ppm", choices=["r50ppm", "hrnet"])
    ap.add_argument("--out", default="gpurun_out/timeline.csv")
    ap.add_argument("--eager", action="store_true", help="profile an eager run instead of the graph replay")
    ap.add_argument("--with-sgd", action="store_true", help="include the FusedSGD step (as bench.py's value arm does)")
    args = ap.parse_args()
    from torch.profiler import profile, ProfilerActivity
    from mit_semseg.engine.program import SegProgram
    dev = torch.device("cuda", 0)
    stride = 8
    if args.net == "r50ppm":
        seg = bench.build_model(dev)
    else:
        import torch.nn as nn
        from mit_semseg.models import ModelBuilder, SegmentationModule
        from mit_semseg.models import hrnet as HR
        torch.manual_seed(304)
        enc, dec = HR.hrnetv2(pretrained=False), ModelBuilder.build_decoder("c1", fc_dim=720, num_class=150)
        stride = 4
        seg = SegmentationModule(enc, dec, nn.NLLLoss(ignore_index=-1), None).to(dev).train()
    feed = bench.synth_batch(2, 512, 512, stride, 304)
    prog = SegProgram(seg, tuple(feed["img_data"].shape), training=True, with_grad=True)
    prog.load_inputs(feed["img_data"].to(dev), feed["seg_label"].to(dev))
    opts = []
    if args.with_sgd:
        opts = bench.make_optimizers(seg, fused=True)
        grads = prog.param_grads()
        for p in seg.parameters():
            p.grad = grads[p]
    if not args.eager:
        prog.capture()

    def step():
        prog.run() if not args.eager else prog.run_eager()
        for o in opts:
            o.step()
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
        step()
        torch.cuda.synchronize()
    with tempfile.NamedTemporaryFile(suffix=".json", delete=False) as f:
        path = f.name
    prof.export_chrome_trace(path)
    tr = json.load(open(path))
    os.unlink(path)
    ks = [e for e in tr["traceEvents"] if e.get("cat") == "kernel" and e.get("ph") == "X"]
    ks.sort(key=lambda e: e["ts"])
    # two identical steps were recorded: keep the second one
    ks = ks[len(ks) // 2:]
    t0 = ks[0]["ts"]
    end = max(e["ts"] + e["dur"] for e in ks)
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    main_stream = collections.Counter(e["args"].get("stream") for e in ks).most_common(1)[0][0]
    with open(args.out, "w") as f:
        f.write("start_us,dur_us,stream,grid,block,regs,smem,name\n")
        for e in ks:
            a = e["args"]
            f.write("%.3f,%.3f,%s,%s,%s,%s,%s,%s\n" % (e["ts"] - t0, e["dur"], a.get("stream"),
                                                      "x".join(str(v) for v in a.get("grid", [])),
                                                      "x".join(str(v) for v in a.get("block", [])),
                                                      a.get("registers per thread"), a.get("shared memory"),
                                                      short(e["name"])))
    fam = collections.OrderedDict()
    for e in ks:
        d = fam.setdefault((short(e["name"]), e["args"].get("stream") == main_stream), [0, 0.0])
        d[0] += 1
        d[1] += e["dur"]
    # union of busy intervals (any stream) and of the main stream alone
    def union(evs):
        tot, cur_s, cur_e = 0.0, None, None
        for e in sorted(evs, key=lambda e: e["ts"]):
            s, t = e["ts"], e["ts"] + e["dur"]
            if cur_e is None or s > cur_e:
                if cur_e is not None:
                    tot += cur_e - cur_s
                cur_s, cur_e = s, t
            else:
                cur_e = max(cur_e, t)
        return tot + (cur_e - cur_s if cur_e is not None else 0.0)
    mk = [e for e in ks if e["args"].get("stream") == main_stream]
    print("step span %.1f us, %d kernels (%d on the main stream %s)" % (end - t0, len(ks), len(mk), main_stream))
    print("busy (any stream) %.1f us; main stream busy %.1f us, idle between its kernels %.1f us" %
          (union(ks), union(mk), (end - t0) - union(mk)))
    gl = sorted(((mk[i + 1]["ts"] - (mk[i]["ts"] + mk[i]["dur"])) for i in range(len(mk) - 1)))
    if gl:
        pos = [g for g in gl if g > 0]
        print("main-stream gaps: median %.2f us, mean %.2f us, >2us: %d, sum of positive gaps %.1f us, overlapped (PDL) pairs %d" %
              (gl[len(gl) // 2], sum(gl) / len(gl), sum(1 for g in gl if g > 2), sum(pos), sum(1 for g in gl if g < 0)))
    print("%-62s %5s %6s %10s %8s" % ("kernel", "main", "calls", "us", "avg"))
    for (name, on_main), (cnt, us) in sorted(fam.items(), key=lambda kv: -kv[1][1])[:40]:
        print("%-62s %5s %6d %10.1f %8.2f" % (name, "main" if on_main else "side", cnt, us, us / cnt))


if __name__ == "__main__":
    main()
