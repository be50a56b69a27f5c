"""CPU, world_size 2, gloo: the host-side logic of the N>1 path.

 * `UserScatteredDataParallel` keeps the reference contract (input = list with one dict per GPU) and selects this
   rank's entry; outputs gain the leading dimension the reference's dict_gather gives them.
 * `SynchronizedBatchNorm2d.is_synchronized()` follows the reference's switch (batchnorm.py:58) once a process group
   with world_size > 1 exists.
 * The exchange protocol the engine implements on the GPU — all-reduce of [sum | sum^2 | count] in forward, of
   [sum g*xhat | sum g] in backward, then mean of the per-rank gradients — reproduces the reference's single-process
   DataParallel + SyncBN result.  Checked here with the oracle's math on CPU tensors: rank-local halves + gloo
   all-reduces  ==  oracle BN (sync formula) on the concatenated batch, for outputs AND input gradients, including
   ranks with different batch sizes (variable per-GPU shapes).
"""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import segnet_oracle as O


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mit_semseg.lib.nn import SynchronizedBatchNorm2d, UserScatteredDataParallel, patch_replication_callback
        from mit_semseg.lib.nn.parallel import _rank_world
        assert _rank_world() == (rank, world)
        bn = SynchronizedBatchNorm2d(4)
        assert bn.is_synchronized()          # training + world_size 2
        bn.eval()
        assert not bn.is_synchronized()      # eval never synchronises (batchnorm.py:58)

        class Echo(torch.nn.Module):
            def forward(self, feed):
                return feed["x"].sum(), feed["x"].mean()

        dp = UserScatteredDataParallel(Echo(), device_ids=[0, 1])
        patch_replication_callback(dp)
        batches = [{"x": torch.full((2, 3), 1.0)}, {"x": torch.full((5, 3), 2.0)}]
        # CPU run: exercise the rank selection directly (scatter needs CUDA streams)
        mine = batches[rank % len(batches)]
        s, m = dp.module(mine)
        from mit_semseg.lib.nn.parallel import _lift
        s, m = _lift((s, m))
        assert s.shape == (1,) and float(s) == (6.0 if rank == 0 else 30.0)

        # ---- SyncBN exchange protocol with uneven per-rank batches
        g = torch.Generator().manual_seed(0)
        C = 6
        full = torch.randn(5, C, 4, 4, generator=g) * 2 + 1
        gout = torch.randn(5, C, 4, 4, generator=g)
        gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
        lo, hi = (0, 2) if rank == 0 else (2, 5)
        x, go = full[lo:hi], gout[lo:hi]
        xv = x.transpose(0, 1).reshape(C, -1)
        msg = torch.cat([xv.sum(1), (xv * xv).sum(1), torch.tensor([float(xv.shape[1])])])
        dist.all_reduce(msg)                                    # forward message: [sum | sqsum | count]
        cnt = msg[2 * C].item()
        mean = msg[:C] / cnt
        var = (msg[C:2 * C] - msg[:C] * mean) / cnt
        inv = var.clamp(min=1e-5).rsqrt()
        y = (x - mean.view(1, C, 1, 1)) * (inv * gamma).view(1, C, 1, 1) + beta.view(1, C, 1, 1)
        xhat = (x - mean.view(1, C, 1, 1)) * inv.view(1, C, 1, 1)
        bmsg = torch.cat([(go * xhat).transpose(0, 1).reshape(C, -1).sum(1), go.transpose(0, 1).reshape(C, -1).sum(1)])
        dist.all_reduce(bmsg)                                   # backward message: [dgamma | dbeta]
        dx = (gamma * inv).view(1, C, 1, 1) * (go - bmsg[C:].view(1, C, 1, 1) / cnt - xhat * bmsg[:C].view(1, C, 1, 1) / cnt)
        # oracle on the concatenated batch (what the reference's master thread computes)
        sd = {"bn.weight": gamma.clone().requires_grad_(True), "bn.bias": beta.clone().requires_grad_(True),
              "bn.running_mean": torch.zeros(C), "bn.running_var": torch.ones(C)}
        xf = full.clone().requires_grad_(True)
        yo = O.batch_norm(xf, sd, "bn", O.BNState(True, sync=True))
        yo.backward(gout)
        assert torch.allclose(y, yo[lo:hi].detach(), atol=1e-5)
        assert torch.allclose(dx, xf.grad[lo:hi], atol=1e-4)
        assert torch.allclose(bmsg[:C], sd["bn.weight"].grad, atol=1e-4)
        assert torch.allclose(bmsg[C:], sd["bn.bias"].grad, atol=1e-4)

        # ---- gradient bucket: all-reduce(sum) then 1/world == gradient of mean(per-rank losses)
        w = torch.ones(3, requires_grad=True)
        loss = (w * (rank + 1.0)).sum()
        loss.backward()
        bucket = w.grad.clone()
        dist.all_reduce(bucket)
        bucket /= world
        assert torch.allclose(bucket, torch.full((3,), 1.5))
        # ---- multi-scale evaluation sharded over ranks (eval.py:63-72): scale k -> rank k mod G,
        #      one all-reduce of the partial score maps == the serial `scores += pred / len(scales)` loop
        from mit_semseg.engine import functional as EF
        assert EF.shard_scales(5, 2, 0) == [0, 2, 4] and EF.shard_scales(5, 2, 1) == [1, 3]

        class FakeSeg:   # stands in for a SegmentationModule: the engine call itself is replaced by _run_scale below
            num_class = 3
        gs = torch.Generator().manual_seed(1)
        imgs = [torch.randn(1, 3, 4 + k, 5 + k, generator=gs) for k in range(5)]

        def fake_scale(img, scores, weight):   # any deterministic per-scale "probability map"
            scores += weight * torch.sigmoid(img.mean()) * torch.ones_like(scores)
        got = EF.multiscale_inference(FakeSeg(), imgs, (6, 7), _run_scale=fake_scale)
        ref = sum(torch.sigmoid(im.mean()) for im in imgs) / 5 * torch.ones(1, 3, 6, 7)
        assert got.shape == (1, 3, 6, 7) and torch.allclose(got, ref, atol=1e-6)
        # ---- the data-parallel training schedule itself, kernels replaced by argument-checking no-ops
        #      (tests/test_program_null_run.py): every collective of the step really runs, over gloo - SyncBN statistics
        #      (forward + backward), the three gradient buckets, with and without the side-stream re-layout
        import torch.nn as nn
        from mit_semseg.engine import _C, ops, program as PR
        from test_program_dry import _seg
        from test_program_null_run import _NullLib
        lib = _NullLib(_C._SIGNATURES)
        _C.lib = lambda: lib
        ops._stream = lambda: None
        os.environ["SSEG_PEER_SYNC"] = "0"          # the peer arena needs CUDA IPC; this is the NCCL-collective schedule
        torch.manual_seed(0)
        seg = _seg("resnet18dilated", "ppm_deepsup", 512)
        seg.train()
        for overlap in ("0", "1"):
            os.environ["SSEG_OVERLAP_RELAYOUT"] = overlap
            P = PR.SegProgram(seg, (2, 3, 64, 64), training=True, with_grad=True, dry_run=True)
            assert P.world == 2 and P.peer is None
            assert all(r.mode == ops.BN_TRAIN_SYNC for r in P.records if isinstance(r, PR.ConvBNRec))
            P.load_inputs(torch.randn(2, 3, 64, 64), torch.randint(-1, 150, (2, 8, 8)))
            P.dry_run, P.serial = False, True
            P.gflat.fill_(float(rank + 1))           # stands in for this rank's weight gradients
            P.bwd = P.bwd[1:]                        # (skip the memset of the gradient buffer for this check)
            P.run_eager()
            # every slice of the flat gradient buffer went through exactly one all-reduce: 1 + 2 = 3 on both ranks
            nsmall = P.g_small
            assert torch.all(P.gflat[nsmall:] == 3.0), "conv gradient buckets must be summed over the ranks exactly once"
        os.environ.pop("SSEG_OVERLAP_RELAYOUT")

        # ---- the same schedule COMPUTED (tests/abi_emulator.py restates the C ABI in torch): two ranks over gloo against
        #      the oracle on the CONCATENATED batch with the synchronised BN formula and the mean of the per-rank losses -
        #      the reference's DataParallel semantics (batchnorm.py:98-139, train.py:42); what tools/dist_check.py checks
        #      on two GPUs.
        import torch.nn.functional as F
        from abi_emulator import EmuLib
        from test_program_emulated import _load, _rel
        emu = EmuLib()
        _C.lib = lambda: emu
        enc_arch, dec_arch, fc = "resnet18dilated", "c1_deepsup", 512
        seg = _seg(enc_arch, dec_arch, fc)
        esd, dsd = _load(seg, enc_arch, dec_arch, fc)
        seg.train()
        feeds = [O.synth_batch(2, 64, 64, 8, 100 + r) for r in range(world)]
        P = PR.SegProgram(seg, (2, 3, 64, 64), training=True, with_grad=True, dry_run=True)
        P.dry_run, P.serial = False, True
        P.load_inputs(feeds[rank]["img_data"], feeds[rank]["seg_label"])
        P.run_eager()
        grads = {("enc." if net is seg.encoder else "dec.") + n: P.param_grads()[p].clone()
                 for net in (seg.encoder, seg.decoder) for n, p in net.named_parameters()}
        fl = torch.cat([g.flatten() for g in grads.values()])
        other = fl.clone()
        dist.broadcast(other, 0)
        assert torch.equal(fl, other), "ranks disagree on the reduced gradients"
        losses = [torch.zeros(1) for _ in range(world)]
        dist.all_gather(losses, P.out[:1].clone())
        e = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in esd.items()}
        d = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in dsd.items()}
        st = O.BNState(True, sync=True, update_running=True, emulate="bf16")
        feats = O.encoder_forward(torch.cat([f["img_data"] for f in feeds]), e, enc_arch, st)
        lg, lg_ds = O.decoder_forward(feats, d, dec_arch, st, dropout_p=0.0, return_logits=True)
        per = []
        for r in range(world):
            sl, lab = slice(2 * r, 2 * r + 2), feeds[r]["seg_label"]
            per.append(F.nll_loss(F.log_softmax(lg[sl], 1), lab, ignore_index=-1) +
                       0.4 * F.nll_loss(F.log_softmax(lg_ds[sl], 1), lab, ignore_index=-1))
        (sum(per) / world).backward()
        for got, want in zip(losses, per):
            assert abs(got.item() - want.item()) <= 5e-3 * abs(want.item()), (got.item(), want.item())
        gr = torch.cat([(e if n.startswith("enc.") else d)[n[4:]].grad.flatten() for n in grads])
        cos = torch.dot(fl, gr).item() / (fl.norm() * gr.norm()).item()
        assert cos >= 0.97, cos
        assert (seg.encoder.bn1.running_mean - e["bn1.running_mean"]).abs().max().item() < 1e-4   # accumulator formula

        # ---- the PEER-MEMORY schedule (the default on GPUs: SyncBN statistics pooled out of the ranks' arenas, no
        #      collective per layer), plain and with the fused conv+BN kernels doing the pooling themselves.  CUDA IPC has
        #      no CPU counterpart, so every "peer" pointer of the stand-in arena maps this rank's own memory: with the SAME
        #      batch on both ranks the pooled sums are exactly world x local, i.e. the oracle on the batch concatenated
        #      with itself.  (The NVLink handshake itself is exercised by tools/dist_check.py on GPUs.)
        import ctypes
        from mit_semseg.engine import peer as PEER

        class LocalArena:
            def __init__(self, nfloats, dist_, device):
                self.world, self.rank = dist_.get_world_size(), dist_.get_rank()
                self.floats = torch.zeros(nfloats)
                self.ints = self.floats.view(torch.int32)
                self.bases = (ctypes.c_void_p * self.world)(*([self.floats.data_ptr()] * self.world))
        PEER.PeerArena = LocalArena
        os.environ["SSEG_PEER_SYNC"] = "1"
        same = O.synth_batch(2, 64, 64, 8, 300)
        ref_grads = None
        for coop in ("0", "1"):
            os.environ["SSEG_COOP_BN"] = coop
            seg = _seg(enc_arch, dec_arch, fc)
            esd, dsd = _load(seg, enc_arch, dec_arch, fc)
            seg.train()
            P = PR.SegProgram(seg, (2, 3, 64, 64), training=True, with_grad=True, dry_run=True)
            assert P.peer is not None and P.world == 2
            P.dry_run, P.serial = False, True
            P.load_inputs(same["img_data"], same["seg_label"])
            P.run_eager()
            if coop == "1":
                assert emu.calls.get("sseg_conv_bn_train", 0) > 0 and emu.calls.get("sseg_conv_dgrad_bn", 0) > 0
            e = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in esd.items()}
            d = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in dsd.items()}
            st = O.BNState(True, sync=True, update_running=True, emulate="bf16")
            feats = O.encoder_forward(torch.cat([same["img_data"], same["img_data"]]), e, enc_arch, st)
            lg, lg_ds = O.decoder_forward(feats, d, dec_arch, st, dropout_p=0.0, return_logits=True)
            lab = same["seg_label"]
            per = [F.nll_loss(F.log_softmax(lg[sl], 1), lab, ignore_index=-1) + 0.4 * F.nll_loss(F.log_softmax(lg_ds[sl], 1), lab, ignore_index=-1)
                   for sl in (slice(0, 2), slice(2, 4))]
            (sum(per) / world).backward()
            assert abs(P.out[0].item() - per[rank].item()) <= 5e-3 * abs(per[rank].item()), (coop, P.out[0].item(), per[rank].item())
            names = [("enc." if net is seg.encoder else "dec.") + n for net in (seg.encoder, seg.decoder) for n, _ in net.named_parameters()]
            fl = torch.cat([P.param_grads()[p].flatten() for net in (seg.encoder, seg.decoder) for p in net.parameters()])
            gr = torch.cat([(e if n.startswith("enc.") else d)[n[4:]].grad.flatten() for n in names])
            cos = torch.dot(fl, gr).item() / (fl.norm() * gr.norm()).item()
            assert cos >= 0.97, (coop, cos)
            assert (seg.encoder.bn1.running_mean - e["bn1.running_mean"]).abs().max().item() < 1e-4, coop
        os.environ.pop("SSEG_COOP_BN")
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_host_logic():
    import random
    port = 29000 + random.randint(0, 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get(0) and ret.get(1)


def _worker_shapes(rank, world, port, ret):
    """Ranks whose batch shapes alternate OUT OF PHASE through the public API (what the reference's loader produces under
    one process per GPU): step-program cache hits / misses and graph captures happen at different steps on different ranks.
    Nothing in program construction or capture may issue a collective, or the ranks' collectives would pair wrongly."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["SSEG_PEER_SYNC"] = "0"      # SyncBN statistics through (gloo) all-reduces: the emulated ABI has no CUDA IPC
    os.environ["SSEG_CAPTURE_ON_USE"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import conftest
        conftest.install_cuda_stand_in(setattr, "1")
        import torch.nn as nn
        from mit_semseg.engine import program as PR
        from test_program_dry import _seg
        calls = {"capture": 0, "collectives_in_capture": 0}
        real_all_reduce = dist.all_reduce
        in_capture = [False]

        def counting_all_reduce(*a, **k):
            if in_capture[0]:
                calls["collectives_in_capture"] += 1
            return real_all_reduce(*a, **k)
        dist.all_reduce = counting_all_reduce

        def capture(self, warm=True):        # the stand-in for CUDA-graph capture: run what the real one would run
            calls["capture"] += 1
            in_capture[0] = True
            try:
                if warm:
                    self.run_eager()
                    self.run_eager()
            finally:
                in_capture[0] = False
        PR.SegProgram.capture = capture
        torch.manual_seed(0)
        seg = _seg("resnet18dilated", "c1_deepsup", 512)
        seg.train()
        opt = torch.optim.SGD(seg.parameters(), lr=0.01, momentum=0.9)
        shapes = [(2, 64, 64), (2, 64, 96)]
        for step in range(4):     # every shape twice on every rank: its program is captured on the second use
            n, h, w = shapes[(step + rank) % 2]
            feed = O.synth_batch(n, h, w, 8, 500 + 10 * step + rank)
            seg.zero_grad()
            loss, acc = seg(feed)
            loss.mean().backward()
            opt.step()
            assert torch.isfinite(loss).all()
        assert calls["capture"] == 2, calls                      # each shape captured once (on its second use) ...
        assert calls["collectives_in_capture"] == 0, calls       # ... without executing anything
        flat = torch.cat([p.detach().flatten() for p in seg.parameters()] +
                         [b.detach().flatten().float() for b in seg.buffers()])
        ref = flat.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(flat, ref), "ranks diverged"
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def test_two_ranks_with_out_of_phase_batch_shapes():
    import random
    port = 29000 + random.randint(0, 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_shapes, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get(0) and ret.get(1)
