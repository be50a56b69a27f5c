"""SURVEY 8(f) row 4, measurement (first run next round, one GPU): the bench network's end-to-end training step fed by
  (a) bench.py's e2e arm as it is: fp32 NCHW image + int64 labels from pinned memory, copied synchronously in front of the step;
  (b) the same fp32 batches through DevicePrefetcher (copy stream, one batch in flight);
  (c) raw batches - uint8 HWC image + uint8 labels, the bytes TrainDataset(raw=True) yields - through DevicePrefetcher with
      the device-side img_transform / segm_transform.
Prints one line per arm: images/s, ms/step, H2D bytes per step.   python tools/input_pipeline_bench.py [--steps 30]"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "semantic-segmentation-pytorch_b200")):
    sys.path.insert(0, p)
import bench  # noqa: E402


def _pin(t):
    try:
        return t.pin_memory()
    except RuntimeError:     # dry run without a CUDA driver (emulated ABI)
        return t


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--crop", type=int, default=bench.CROP)
    ap.add_argument("--dec", default=None, help="another decoder (with --fc) for a quick dry run on the emulated ABI")
    ap.add_argument("--fc", type=int, default=None)
    args = ap.parse_args()
    if args.dec:
        bench.DEC_ARCH, bench.FC_DIM = args.dec, args.fc
    from mit_semseg.engine.prefetch import DevicePrefetcher
    dev = torch.device("cuda", 0)
    seg = bench.build_model(dev)
    opts = bench.make_optimizers(seg)
    n, c = bench.BATCH, args.crop
    g = torch.Generator().manual_seed(7)
    pool = []
    for i in range(4):    # a few distinct batches of the bench shape
        u8 = torch.randint(0, 256, (n, c, c, 3), generator=g).to(torch.uint8)
        seg_u8 = torch.randint(0, bench.NUM_CLASS + 1, (n, c // bench.LABEL_STRIDE, c // bench.LABEL_STRIDE), generator=g).to(torch.uint8)
        valid = torch.tensor([[c, c]] * n, dtype=torch.int32)
        f32 = ((u8.float() / 255.).permute(0, 3, 1, 2) - torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)) / torch.tensor(
            [0.229, 0.224, 0.225]).view(1, 3, 1, 1)
        pool.append(({"img_u8": u8, "seg_u8": seg_u8, "valid_hw": valid, "segm_downsampling_rate": bench.LABEL_STRIDE},
                     {"img_data": _pin(f32.contiguous()), "seg_label": _pin(seg_u8.long() - 1)}))

    def batches(kind, count):
        for i in range(count):
            yield pool[i % len(pool)][0 if kind == "raw" else 1]

    def train_on(feed):
        seg.zero_grad()
        loss, acc = seg(feed)
        loss = loss.mean()
        loss.backward()
        for o in opts:
            o.step()
        return loss.item()

    def run(arm):
        total = args.warmup + args.steps
        if arm == "sync-f32":
            src = ({k: v.to(dev, non_blocking=True) for k, v in b.items()} for b in batches("f32", total))
            nbytes = sum(v.numel() * v.element_size() for v in pool[0][1].values())
        else:
            pf = DevicePrefetcher(batches("raw" if arm == "prefetch-u8" else "f32", total), device=dev)
            src = iter(pf)
        t0 = None
        for i, feed in enumerate(src):
            if i == args.warmup:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
            last = train_on(feed)
            if arm != "sync-f32":
                nbytes = pf.h2d_bytes
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print("%-12s %8.1f images/s  %7.3f ms/step  H2D %9d B/step  loss %.4f" % (arm, n * args.steps / dt, dt / args.steps * 1e3, nbytes, last))

    for arm in ("sync-f32", "prefetch-f32", "prefetch-u8"):
        run(arm)


if __name__ == "__main__":
    main()
