#!/usr/bin/env python
"""Where the end-to-end step (bench.py's `e2e` arm: public API, pinned host inputs, loss read-back every step) spends its
time beyond the CUDA-graph replay: host wall-clock per phase of the loop body (the GPU runs behind the host, so a phase is
on the critical path only when the GPU is idle during it: everything between `loss.item()` returning and the graph's first
kernel) and the device-side span of one step (CUDA events around H2D copy .. SGD).

    python tools/e2e_breakdown.py [--steps 50]"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "semantic-segmentation-pytorch_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--no-zero-grad", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    seg = bench.build_model(dev)
    opts = bench.make_optimizers(seg, fused=True)
    f = bench.synth_batch(bench.BATCH, bench.CROP, bench.CROP, bench.LABEL_STRIDE, 304)
    feed = {k: v.pin_memory() for k, v in f.items()}
    names = ["zero_grad", "seg(feed)", "loss.mean", "backward", "opt.step", "item"]
    acc = dict.fromkeys(names, 0.0)
    dev_ms = 0.0

    def step(record):
        nonlocal dev_ms
        t = [time.perf_counter()]
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        if not args.no_zero_grad:
            seg.zero_grad()
        t.append(time.perf_counter())
        loss, _ = seg(feed)
        t.append(time.perf_counter())
        loss = loss.mean()
        t.append(time.perf_counter())
        loss.backward()
        t.append(time.perf_counter())
        for o in opts:
            o.step()
        t.append(time.perf_counter())
        e1.record()
        v = loss.item()
        t.append(time.perf_counter())
        if record:
            for i, n in enumerate(names):
                acc[n] += t[i + 1] - t[i]
            dev_ms += e0.elapsed_time(e1)
        return v

    for _ in range(8):
        step(False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(True)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / args.steps * 1e3
    print("e2e step %.3f ms wall; device span (first enqueue .. SGD done) %.3f ms" % (wall, dev_ms / args.steps))
    for n in names:
        print("  host %-10s %7.3f ms" % (n, acc[n] / args.steps * 1e3))
    print("  (item() = waiting for the GPU; GPU idle per step ~ wall - device busy; graph replay alone: see bench value)")


if __name__ == "__main__":
    main()
