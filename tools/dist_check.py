#!/usr/bin/env python
"""Multi-GPU check (run under torchrun, one rank per GPU):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py
 1. SyncBN + gradient bucket over NCCL: every rank ends with the same gradients; they match the oracle run on the
    CONCATENATED batch with the synchronised BN formula and the mean of per-rank losses (the reference's DataParallel
    semantics: batchnorm.py:98-139, train.py:42).
 2. The same program captured as one CUDA graph (NCCL calls inside) reproduces the eager result.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "semantic-segmentation-pytorch_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn as nn  # noqa: E402
import torch.nn.functional as F  # noqa: E402


def syncbn_module_test(rank, world):
    """lib/nn/modules/tests/test_sync_batchnorm.py:98-108 (testSyncBatchNorm2DSyncTrain) with the replicas as processes:
    each rank feeds its slice of the batch through SynchronizedBatchNorm2d; outputs and input gradients must equal
    nn.BatchNorm2d on the WHOLE batch (one bf16 rounding of the value scale), the parameter gradients must sum to its."""
    from mit_semseg.lib.nn import SynchronizedBatchNorm2d
    g = torch.Generator().manual_seed(0)
    N, C = 16, 16
    full = torch.rand(N, C, 16, 16, generator=g).bfloat16().float()
    gout = torch.randn(N, C, 16, 16, generator=g).bfloat16().float()
    gamma, beta = 0.5 + torch.rand(C, generator=g), torch.randn(C, generator=g) * 0.1
    bn, sbn = nn.BatchNorm2d(C).cuda(), SynchronizedBatchNorm2d(C).cuda()
    for m in (bn, sbn):
        m.weight.data.copy_(gamma)
        m.bias.data.copy_(beta)
    assert sbn.is_synchronized()
    per = N // world
    lo, hi = rank * per, (rank + 1) * per
    x = full[lo:hi].cuda().requires_grad_(True)
    out = sbn(x)
    (out * gout[lo:hi].cuda()).sum().backward()
    xr = full.cuda().requires_grad_(True)
    ref = bn(xr)
    (ref * gout.cuda()).sum().backward()

    def close(a, b, what, rel=2 ** -8):
        tol = rel * max(b.abs().max().item(), 1e-3)
        err = (a - b).abs().max().item()
        assert err <= tol, "rank %d %s: %.3e > %.3e" % (rank, what, err, tol)
    close(out.detach(), ref.detach()[lo:hi], "output")
    close(x.grad, xr.grad[lo:hi], "input gradient")
    gw, gb = sbn.weight.grad.clone(), sbn.bias.grad.clone()
    dist.all_reduce(gw)
    dist.all_reduce(gb)
    close(gw, bn.weight.grad, "weight gradient (summed over ranks)", rel=2 ** -7)
    close(gb, bn.bias.grad, "bias gradient (summed over ranks)", rel=2 ** -7)
    ok = torch.ones(1, device="cuda")
    dist.all_reduce(ok)
    if rank == 0:
        print("SYNCBN-TEST OK (world %d)" % world, flush=True)


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if "--syncbn-test" in sys.argv:
        syncbn_module_test(rank, world)
        import bench
        bench.shutdown_distributed()
        return
    from test_gpu_e2e import _build
    from mit_semseg.engine.program import SegProgram
    from mit_semseg.engine import ops
    from oracle import segnet_oracle as O
    enc_arch, dec_arch, fc = "resnet18dilated", "c1_deepsup", 512
    seg, esd, dsd, ds = _build(enc_arch, dec_arch, fc, residual_gain=0.25)
    seg.cuda().train()
    n, hw = 2, 96
    feeds = [O.synth_batch(n, hw, hw, 8, 100 + r) for r in range(world)]
    mine = feeds[rank]
    prog = SegProgram(seg, tuple(mine["img_data"].shape), training=True, with_grad=True)
    assert prog.world == world and prog._bn_mode(next(iter(prog.bns.values()))) == ops.BN_TRAIN_SYNC
    prog.load_inputs(mine["img_data"].cuda(), mine["seg_label"].cuda())
    prog.run_eager()
    torch.cuda.synchronize()
    loss_eager = prog.out[0].item()
    grads = {name: prog.param_grads()[p].detach().clone() for name, p in
             list(seg.encoder.named_parameters(prefix="enc")) + list(seg.decoder.named_parameters(prefix="dec"))}
    rm_engine = seg.encoder.bn1.running_mean.clone()
    # cross-rank consistency
    flat = torch.cat([g.flatten() for g in grads.values()])
    ref = flat.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(flat, ref), "ranks disagree on the reduced gradients"
    losses = [torch.zeros(1, device="cuda") for _ in range(world)]
    dist.all_gather(losses, torch.tensor([loss_eager], device="cuda"))
    if rank == 0:
        e = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in esd.items()}
        d = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in dsd.items()}
        st = O.BNState(True, sync=True, update_running=True, emulate="bf16")
        img = torch.cat([f["img_data"] for f in feeds])
        feats = O.encoder_forward(img, e, enc_arch, st)
        lg, lg_ds = O.decoder_forward(feats, d, dec_arch, st, dropout_p=0.0, return_logits=True)
        per = []
        for r in range(world):
            sl = slice(r * n, (r + 1) * n)
            lab = feeds[r]["seg_label"]
            per.append(F.nll_loss(F.log_softmax(lg[sl], 1), lab, ignore_index=-1) +
                       0.4 * F.nll_loss(F.log_softmax(lg_ds[sl], 1), lab, ignore_index=-1))
        total = sum(per) / world
        total.backward()
        num = den = dot = 0.0
        worst = (0, "")
        for name, g in grads.items():
            sd = e if name.startswith("enc.") else d
            gr = sd[name[4:]].grad.double().flatten()
            gg = g.cpu().double().flatten()
            dot += torch.dot(gg, gr).item(); num += torch.dot(gg, gg).item(); den += torch.dot(gr, gr).item()
            rel = ((gg - gr).norm() / (gr.norm() + 1e-30)).item()
            worst = max(worst, (rel, name))
        cos = dot / (num ** 0.5 * den ** 0.5)
        print("per-rank loss engine %s oracle %s" % ([round(l.item(), 5) for l in losses], [round(p.item(), 5) for p in per]))
        print("gradient cosine vs oracle(concatenated batch, sync BN) %.4f  norm ratio %.4f  worst rel %.3f (%s)" %
              (cos, (num / den) ** 0.5, worst[0], worst[1]))
        rm_o = e["bn1.running_mean"]
        print("running_mean (sync accumulator formula) max |diff| %.3e" % (rm_engine.cpu() - rm_o).abs().max().item())
        for l, p in zip(losses, per):
            assert abs(l.item() - p.item()) <= 5e-3 * abs(p.item())
        assert cos >= 0.9 and (rm_engine.cpu() - rm_o).abs().max().item() < 1e-3
    # ---- CUDA graph with NCCL inside
    dist.barrier()
    prog.capture()
    prog.run()
    torch.cuda.synchronize()
    loss_graph = prog.out[0].item()
    g2 = prog.param_grads()[seg.decoder.conv_last.weight]
    cosg = F.cosine_similarity(g2.flatten(), grads["dec.conv_last.weight"].flatten(), dim=0).item()
    print("rank %d: eager loss %.5f graph loss %.5f  head-grad cosine %.4f" % (rank, loss_eager, loss_graph, cosg))
    assert abs(loss_graph - loss_eager) <= 2e-3 * abs(loss_eager) and cosg > 0.95
    dist.barrier()
    if rank == 0:
        print("DIST_CHECK_OK", flush=True)
    import bench
    prog.graph = None
    bench.shutdown_distributed()


if __name__ == "__main__":
    main()
