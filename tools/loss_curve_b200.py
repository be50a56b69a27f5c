#!/usr/bin/env python
"""Training loss curve of BASELINE config[2] (ResNet50dilated + PPM_deepsup, 2 x 3 x 512 x 512 per step) ON THE B200:
the engine through the public API (SegmentationModule(feed) -> loss.backward() -> torch.optim.SGD, train.py:41-48) next to
the reference's arithmetic on the same GPU - the oracle port (same torch ops in the reference's order) executed by stock
PyTorch + cuDNN in fp32 (TF32 off) and, for the run-to-run noise band, a second fp32 run whose batches are visited in the
same order but whose cuDNN algorithms are non-deterministic, plus a bf16-autocast run.

Same initial weights (reference initialisers, seed 304), same `--batches` synthetic batches visited cyclically, same
optimiser and poly learning-rate schedule (train.py:130-139), Dropout2d off on both sides (its draws would come from
different RNG streams).   python tools/loss_curve_b200.py [--steps 200] [--batches 8] > profiles/r2_loss_curve_b200.txt"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "semantic-segmentation-pytorch_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--batches", type=int, default=8)
    ap.add_argument("--lr", type=float, default=0.02)
    args = ap.parse_args()
    from oracle import segnet_oracle as O
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    seg = bench.build_model(dev)
    for m in seg.modules():
        if isinstance(m, nn.Dropout2d):
            m.p = 0.0
    # the oracle's state dicts = the engine's initial weights
    def clone_sd(net):
        return {k: v.detach().clone().contiguous().requires_grad_(v.is_floating_point() and "running" not in k and "num_batches" not in k)
                for k, v in net.state_dict().items()}
    arms = {}
    for name in ("oracle fp32 (a)", "oracle fp32 (b)", "oracle bf16 autocast"):
        e, d = clone_sd(seg.encoder), clone_sd(seg.decoder)
        params = [v for v in list(e.values()) + list(d.values()) if v.requires_grad]
        arms[name] = (e, d, torch.optim.SGD(params, lr=args.lr, momentum=bench.MOMENTUM, weight_decay=bench.WD))
    # train.py's two optimisers with its parameter groups (weight decay on conv weights only); the oracle arms use one
    # optimiser over all parameters with the same hyper-parameters on conv weights - BN / bias decay differs
    opts = bench.make_optimizers(seg)
    for name, (e, d, opt) in arms.items():   # same grouping for the oracle arms
        decay = [v for k, v in list(e.items()) + list(d.items()) if v.requires_grad and v.dim() == 4]
        rest = [v for k, v in list(e.items()) + list(d.items()) if v.requires_grad and v.dim() != 4]
        arms[name] = (e, d, torch.optim.SGD([dict(params=decay), dict(params=rest, weight_decay=0.0)], lr=args.lr,
                                            momentum=bench.MOMENTUM, weight_decay=bench.WD))
    feeds = [bench.synth_batch(bench.BATCH, bench.CROP, bench.CROP, bench.LABEL_STRIDE, 900 + i) for i in range(args.batches)]
    feeds_dev = [{k: v.to(dev) for k, v in f.items()} for f in feeds]
    feeds_pin = [{k: v.pin_memory() for k, v in f.items()} for f in feeds]
    print("# step  engine(B200 kernels, bf16 storage)  " + "  ".join(arms))
    for step in range(args.steps):
        lr = args.lr * (1.0 - step / float(args.steps)) ** 0.9
        for o in opts + [a[2] for a in arms.values()]:
            for grp in o.param_groups:
                grp["lr"] = lr
        i = step % args.batches
        seg.zero_grad()
        loss, acc = seg(feeds_pin[i])
        loss.mean().backward()
        for o in opts:
            o.step()
        row = [loss.item()]
        for name, (e, d, opt) in arms.items():
            torch.backends.cudnn.benchmark = name.endswith("(b)")   # a different algorithm choice: the run-to-run noise band
            opt.zero_grad()
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled="autocast" in name):
                l, _ = O.segmentation_forward(feeds_dev[i], e, d, bench.ENC_ARCH, bench.DEC_ARCH, O.BNState(True), 0.4,
                                              dropout_p=0.0)
            l.backward()
            opt.step()
            row.append(l.item())
        if step % 5 == 0 or step == args.steps - 1:
            print("%4d  " % step + "  ".join("%8.4f" % v for v in row), flush=True)


if __name__ == "__main__":
    main()
