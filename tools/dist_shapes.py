#!/usr/bin/env python
"""Two (or more) ranks whose batch shapes ALTERNATE OUT OF PHASE through the public API - what the reference's loader does
under one process per GPU (every per-GPU batch draws its own short size / aspect bucket). Rank r sees shape A on even
steps and shape B on odd ones when r is even, the other way round when r is odd: program-cache hits and misses, eager runs
and graph captures therefore happen at different steps on different ranks. Passes when every step completes, the losses are
finite, and after the last step all ranks hold bit-identical weights (SyncBN statistics and gradient buckets met correctly
at every step).   python -m torch.distributed.run --nproc-per-node 2 tools/dist_shapes.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "semantic-segmentation-pytorch_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn as nn  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import bench
    from mit_semseg.models import ModelBuilder, SegmentationModule
    from mit_semseg.models import models as M, resnet as R
    torch.manual_seed(304)
    enc = M.ResnetDilated(R.resnet18(pretrained=False), 8)
    dec = ModelBuilder.build_decoder("ppm_deepsup", fc_dim=512, num_class=150)
    seg = SegmentationModule(enc, dec, nn.NLLLoss(ignore_index=-1), 0.4).cuda().train()
    opts = [torch.optim.SGD(bench.group_weight(net), lr=0.02, momentum=0.9, weight_decay=1e-4) for net in (enc, dec)]
    shapes = [(2, 128, 160), (2, 192, 128)]
    losses = []
    for step in range(10):
        n, h, w = shapes[(step + rank) % 2]
        feed = bench.synth_batch(n, h, w, 8, 1000 + 17 * step + rank)
        feed = {k: v.pin_memory() for k, v in feed.items()}
        seg.zero_grad()
        loss, acc = seg(feed)          # host tensors straight into the public call
        loss.mean().backward()
        for o in opts:
            o.step()
        losses.append(loss.item())
    assert all(l == l and abs(l) < 1e4 for l in losses), losses
    flat = torch.cat([p.detach().flatten() for p in seg.parameters()] + [b.detach().flatten().float() for b in seg.buffers()])
    ref = flat.clone()
    dist.broadcast(ref, 0)
    same = torch.equal(flat, ref)
    ok = torch.tensor([1.0 if same else 0.0], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    progs = seg.__dict__.get("_b200_programs", {})
    print("rank %d: losses %s ... %s; programs cached %d, captured %d; weights identical across ranks: %s" %
          (rank, ["%.3f" % l for l in losses[:2]], ["%.3f" % l for l in losses[-2:]], len(progs),
           sum(1 for q in progs.values() if getattr(q, "graph", None) is not None), same), flush=True)
    assert ok.item() == 1.0, "ranks diverged"
    if rank == 0:
        print("DIST_SHAPES_OK", flush=True)
    for q in progs.values():
        q.graph = None
    bench.shutdown_distributed()


if __name__ == "__main__":
    main()
