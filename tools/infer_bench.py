#!/usr/bin/env python
"""Inference timing of the configurations BASELINE.json lists besides the training bench (GPU box):
  config 2: ResNet18dilated + PPM_deepsup, batch 8 at 512x512           python tools/infer_bench.py --net r18ppm
  config 5: HRNetV2-W48 + C1, one image at 5 scales (short side 300..600) python tools/infer_bench.py --net hrnet --multiscale
Prints ms per call (CUDA events, after warm-up; the programs are captured into CUDA graphs on their second use).
Environment switches under test: SSEG_FOLD_BN_EVAL=1 (BN folded into the conv epilogue)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "semantic-segmentation-pytorch_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--net", default="r18ppm", choices=["r18ppm", "r50ppm", "hrnet"])
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--multiscale", action="store_true", help="one image, short sides 300/375/450/525/600 padded to x32")
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    from mit_semseg.engine import functional as EF
    from mit_semseg.models import ModelBuilder, SegmentationModule
    from mit_semseg.models import hrnet as HR, models as M, resnet as R
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:   # under torch.distributed.run the scales of --multiscale are sharded scale k -> rank k mod world
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.manual_seed(304)
    if args.net == "hrnet":
        enc, dec = HR.hrnetv2(pretrained=False), ModelBuilder.build_decoder("c1", fc_dim=720, num_class=150, use_softmax=True)
    else:
        base, fc = ("resnet18", 512) if args.net == "r18ppm" else ("resnet50", 2048)
        enc = M.ResnetDilated(R.__dict__[base](pretrained=False), 8)
        dec = ModelBuilder.build_decoder("ppm_deepsup", fc_dim=fc, num_class=150, use_softmax=True)
    seg = SegmentationModule(enc, dec, nn.NLLLoss(ignore_index=-1)).to(dev).eval()
    g = torch.Generator().manual_seed(1)
    if args.multiscale:
        # eval.py / dataset.py:ValDataset: original 512 x 683 image, short side -> s, long side capped at 1000, padded to x32
        sizes = []
        for s in (300, 375, 450, 525, 600):
            scale = min(s / 512.0, 1000 / 683.0)
            h, w = int(512 * scale), int(683 * scale)
            sizes.append(((h + 31) // 32 * 32, (w + 31) // 32 * 32))
        imgs = [torch.randn(1, 3, h, w, generator=g).to(dev) for h, w in sizes]
        seg_size = (512, 683)
        call = lambda: EF.multiscale_inference(seg, imgs, seg_size)   # noqa: E731
        what = "multi-scale %s -> %s" % (sizes, seg_size)
    else:
        img = torch.randn(args.batch, 3, args.size, args.size, generator=g).to(dev)
        call = lambda: seg({"img_data": img}, segSize=(args.size, args.size))   # noqa: E731
        what = "%d x 3 x %d x %d" % (args.batch, args.size, args.size)
    with torch.no_grad():
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            call()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    units = 1 if args.multiscale else args.batch
    if rank == 0:
        print("%s %s: %.3f ms / call, %.1f images/s (fold_bn_eval=%s, %d GPU%s)" % (
            args.net, what, ms, units / ms * 1e3, os.environ.get("SSEG_FOLD_BN_EVAL", "0"), world, "s" if world > 1 else ""))
    if world > 1:
        for q in seg.__dict__.get("_b200_programs", {}).values():
            q.graph = None
        import bench
        bench.shutdown_distributed()


if __name__ == "__main__":
    main()
