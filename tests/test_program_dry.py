"""CPU: the host-side schedule builder (engine/program.py) in dry-run mode — module tree + input shape -> records,
buffers and closures, without a GPU and without executing anything (a dry-run program refuses to run: there is no CPU
execution path).  Checks the wiring invariants the backward pass relies on: every convolution gets exactly one
weight-gradient task, every applied activation receives a gradient buffer of its own (8-padded) shape."""
import pytest
import torch
import torch.nn as nn


@pytest.fixture(autouse=True)
def _default_switches(monkeypatch):
    for name in ("SSEG_BRANCH_STREAMS", "SSEG_FOLD_BN_EVAL", "SSEG_OVERLAP_RELAYOUT"):
        monkeypatch.delenv(name, raising=False)


def _seg(enc_arch, dec_arch, fc):
    from mit_semseg.models import ModelBuilder, SegmentationModule
    from mit_semseg.models import hrnet as HR, models as M, resnet as R
    if enc_arch == "hrnetv2":
        enc = HR.hrnetv2(pretrained=False)
    elif enc_arch == "mobilenetv2dilated":
        from mit_semseg.models import mobilenet as MB
        enc = M.MobileNetV2Dilated(MB.mobilenetv2(pretrained=False), dilate_scale=8)
    else:
        dil = enc_arch.endswith("dilated")
        net = R.__dict__[enc_arch.replace("dilated", "")](pretrained=False)
        enc = M.ResnetDilated(net, 8) if dil else M.Resnet(net)
    dec = ModelBuilder.build_decoder(dec_arch, fc_dim=fc, num_class=150)
    ds = 0.4 if dec_arch.endswith("deepsup") else None
    return SegmentationModule(enc, dec, nn.NLLLoss(ignore_index=-1), ds)


@pytest.mark.parametrize("enc,dec,fc,hw,nsum", [("resnet50dilated", "ppm_deepsup", 2048, 64, 0),
                                                 ("resnet50", "upernet_lite", 2048, 64, 0),
                                                 ("hrnetv2", "c1", 720, 64, 26)])
def test_training_schedule_wiring(enc, dec, fc, hw, nsum, monkeypatch):
    from mit_semseg.engine import program as PR
    seg = _seg(enc, dec, fc)
    seg.train()
    side = [0]
    orig = PR.SegProgram.on_side

    def counting(self, fn):
        side[0] += 1
        return orig(self, fn)
    monkeypatch.setattr(PR.SegProgram, "on_side", counting)
    P = PR.SegProgram(seg, (2, 3, hw, hw), training=True, with_grad=True, dry_run=True)
    nconv = sum(isinstance(m, nn.Conv2d) for m in seg.modules())
    # one weight-gradient task per convolution (stem and classifiers included); + the data-gradient operand prep when that
    # runs on the side stream (SSEG_SPLIT_PREP=1, off by default: measured slower)
    assert side[0] == nconv + (1 if P.split_prep else 0)
    assert sum(isinstance(r, PR.SumRec) for r in P.records) == nsum
    for r in P.records:
        a = getattr(r, "a", None)
        if a is None:
            continue
        if isinstance(r, (PR.ConvBNRec, PR.SumRec, PR.StemRec)):
            assert a.g is not None, type(r)
        if a.g is not None:
            assert a.g.shape[:3] == a.tp.shape[:3] and a.g.shape[3] in (a.t.shape[3], a.tp.shape[3])
    grads = P.param_grads()
    assert all(p in grads and grads[p].shape == p.shape for p in seg.parameters())
    with pytest.raises(RuntimeError, match="dry-run"):
        P.run_eager()


def test_hrnet_c1_hidden_layer_is_stored_8_padded():
    """fc_dim 720 -> C1 hidden width 180 (models/models.py:353): storage is 184 wide with zero pad channels, the BN
    vectors the 8-channel-vectorised kernels read are 184 long, parameter gradients keep the parameter's 180."""
    from mit_semseg.engine import program as PR
    seg = _seg("hrnetv2", "c1", 720)
    seg.train()
    P = PR.SegProgram(seg, (1, 3, 32, 32), training=True, with_grad=True, dry_run=True)
    rec = [r for r in P.records if isinstance(r, PR.ConvBNRec) and r.cw.mod is seg.decoder.cbr[0]][0]
    assert rec.y.shape[3] == 184 and rec.a.t.shape[3] == 180 and rec.a.tp.shape[3] == 184 and rec.a.g.shape[3] == 184
    assert rec.bns.scale.numel() == 184 and P.param_grads()[seg.decoder.cbr[1].weight].numel() == 180
    assert len(rec.xs) == 4 and [x.t.shape[3] for x in rec.xs] == [48, 96, 192, 384]   # virtual concat, never materialised
    cls = [r for r in P.records if isinstance(r, PR.ClassifierRec)][0]
    assert cls.cw.wf.shape == (150, 180) and cls.cw.wf.stride(0) == 184              # 16-byte row pitch for TMA


def test_inference_schedule_builds_for_every_encoder():
    from mit_semseg.engine import program as PR
    for enc, dec, fc in (("resnet18dilated", "ppm_deepsup", 512), ("hrnetv2", "c1", 720)):
        seg = _seg(enc, dec, fc)
        seg.eval()
        P = PR.SegProgram(seg, (1, 3, 64, 96), training=False, with_grad=False, seg_size=(64, 96), dry_run=True)
        assert P.probs.shape == (1, 150, 64, 96) and not P.bwd


def test_multiscale_head_accumulates_into_a_shared_score_map():
    """eval.py:72 `scores = scores + scores_tmp / len(imgSizes)`: each scale's program writes into the same buffer."""
    from mit_semseg.engine import program as PR
    seg = _seg("resnet18dilated", "ppm_deepsup", 512)
    seg.eval()
    scores = torch.zeros(1, 150, 64, 96)
    progs = [PR.SegProgram(seg, (1, 3, h, w), training=False, with_grad=False, seg_size=(64, 96), dry_run=True,
                           head_out=scores, head_weight=0.5) for h, w in ((64, 96), (96, 128))]
    assert all(p.probs is scores for p in progs)
    with pytest.raises(AssertionError):
        PR.SegProgram(seg, (1, 3, 64, 96), training=False, with_grad=False, seg_size=(32, 32), dry_run=True, head_out=scores)


@pytest.mark.parametrize("enc,dec,fc,nfork_fwd", [("resnet18dilated", "ppm_deepsup", 512, 4), ("hrnetv2", "c1", 720, 2 + 4 * 3 + 3 * 4)])
def test_branch_streams_fork_and_join_symmetrically(enc, dec, fc, nfork_fwd, monkeypatch):
    """SSEG_BRANCH_STREAMS=1 (opt-in): the pyramid branches / HRNet branches get fork + join closures; every fork is
    joined before the schedule ends, in the forward and in the backward list."""
    from mit_semseg.engine import program as PR
    monkeypatch.setenv("SSEG_BRANCH_STREAMS", "1")
    counts = {"fork": 0, "join": 0}
    for name in ("_fork", "_join"):
        orig = getattr(PR.SegProgram, name)

        def counting(self, k, _orig=orig, _name=name[1:]):
            counts[_name] += 1
            return _orig(self, k)
        monkeypatch.setattr(PR.SegProgram, name, counting)
    seg = _seg(enc, dec, fc)
    seg.train()
    P = PR.SegProgram(seg, (2, 3, 64, 64), training=True, with_grad=True, dry_run=True)
    assert not P._open_branches
    nrows = 26 if enc == "hrnetv2" else 0    # exchange-unit rows branch in the forward schedule only
    assert counts["fork"] == counts["join"] == 2 * nfork_fwd + nrows, counts   # forward + backward
    monkeypatch.setenv("SSEG_BRANCH_STREAMS", "0")
    Q = PR.SegProgram(seg, (2, 3, 64, 64), training=True, with_grad=True, dry_run=True)
    assert len(P.fwd) == len(Q.fwd) + 2 * (nfork_fwd + nrows) and len(P.bwd) == len(Q.bwd) + 2 * nfork_fwd


def test_folded_eval_bn_inference_schedule(monkeypatch):
    """SSEG_FOLD_BN_EVAL=1 (opt-in): inference programs run conv+BN(+shortcut)+ReLU as one launch per layer - the raw conv
    outputs are not even allocated and one launch per BN layer (the apply pass) disappears from the schedule."""
    from mit_semseg.engine import program as PR
    for enc, dec, fc in (("resnet18dilated", "ppm_deepsup", 512), ("resnet50", "upernet", 2048), ("hrnetv2", "c1", 720)):
        seg = _seg(enc, dec, fc)
        seg.eval()
        monkeypatch.setenv("SSEG_FOLD_BN_EVAL", "0")
        P0 = PR.SegProgram(seg, (1, 3, 64, 96), training=False, with_grad=False, seg_size=(64, 96), dry_run=True)
        monkeypatch.setenv("SSEG_FOLD_BN_EVAL", "1")
        P1 = PR.SegProgram(seg, (1, 3, 64, 96), training=False, with_grad=False, seg_size=(64, 96), dry_run=True)
        recs = [r for r in P1.records if isinstance(r, PR.ConvBNRec)]
        assert recs and all(r.folded for r in recs)
        assert all(r.y is None for r in recs if r.apply)
        applied = sum(1 for r in P0.records if isinstance(r, PR.ConvBNRec) and r.apply)
        assert len(P0.fwd) - len(P1.fwd) == applied
        # training programs never fold
        seg.train()
        P2 = PR.SegProgram(seg, (2, 3, 64, 64), training=True, with_grad=True, dry_run=True)
        assert not any(r.folded for r in P2.records if isinstance(r, PR.ConvBNRec))


def test_overlapped_relayout_schedule(monkeypatch):
    """SSEG_OVERLAP_RELAYOUT=1 (opt-in): forward operands of the late layers + all data-gradient operands are produced on
    the side stream (one wait before the first late convolution), gradient re-layout follows each bucket's GEMMs."""
    from mit_semseg.engine import program as PR
    seg = _seg("resnet50dilated", "ppm_deepsup", 2048)
    seg.train()
    Q = PR.SegProgram(seg, (2, 3, 64, 64), training=True, with_grad=True, dry_run=True)
    monkeypatch.setenv("SSEG_OVERLAP_RELAYOUT", "1")
    P = PR.SegProgram(seg, (2, 3, 64, 64), training=True, with_grad=True, dry_run=True)
    t_early, t_late, t_dgrad = P._prep_tables
    assert t_early.n + t_late.n == t_dgrad.n == len(P.convs) - 1            # every conv but the 3-channel stem conv
    early = [c for c in P.convs.values() if c.I != 3 and id(c) not in P._late_convs]
    assert sum(c.O * c.T * c.I for c in early) <= 2_000_000 and not P._late_pending
    assert len(P.fwd) == len(Q.fwd) + 2      # side-stream prep + one wait
    assert len(P.bwd) == len(Q.bwd) + 3      # join for the dgrad operands, 2 bucket closes
    monkeypatch.delenv("SSEG_OVERLAP_RELAYOUT")
    monkeypatch.setenv("SSEG_SPLIT_PREP", "1")   # only the data-gradient operands on the side stream (thin grid)
    R = PR.SegProgram(seg, (2, 3, 64, 64), training=True, with_grad=True, dry_run=True)
    assert R.split_prep and len(R.fwd) == len(Q.fwd) + 1 and len(R.bwd) == len(Q.bwd) + 1
    assert set(P.param_grads()) == set(Q.param_grads())


def test_fused_conv_bn_schedule(monkeypatch):
    """SSEG_COOP_BN=1 (opt-in): single-GPU train-mode layers whose tiles fit the SMs' tensor memory run conv + statistics +
    normalise + shortcut + ReLU as ONE launch; projection shortcuts, the 256x256 stem conv3 and layer4's 2048-channel
    convolutions keep the three-kernel sequence. The backward schedule is unchanged."""
    from mit_semseg.engine import program as PR
    seg = _seg("resnet50dilated", "ppm_deepsup", 2048)
    seg.train()
    Q = PR.SegProgram(seg, (2, 3, 512, 512), training=True, with_grad=True, dry_run=True)
    monkeypatch.setenv("SSEG_COOP_BN", "1")
    P = PR.SegProgram(seg, (2, 3, 512, 512), training=True, with_grad=True, dry_run=True)
    recs = [r for r in P.records if isinstance(r, PR.ConvBNRec)]
    coop = [r for r in recs if r.coop]
    assert len(recs) == 60 and len(coop) == 52
    assert all(r.apply and r.y is not None for r in coop)           # y is kept: the backward pass reads it
    assert not any(r.coop for r in recs if r.cw.O == 2048)
    assert len(Q.fwd) - len(P.fwd) == 2 * len(coop)
    # backward: producers with a single consumer get their whole BN backward from the consumer's data-gradient kernel
    pre = [r for r in P.records if isinstance(r, (PR.ConvBNRec, PR.StemRec)) and r.dy_pre is not None]
    assert len(pre) >= 25 and all(r.fused for r in pre) and len(Q.bwd) - len(P.bwd) == len(pre)
    assert any(isinstance(r, PR.StemRec) for r in pre)   # the stem's BN backward rides in conv2's data-gradient kernel too
    # every BN of a fused layer has a barrier counter slot that the per-step statistics reset zeroes
    assert all(r.bns.counter is not None and r.bns.counter.numel() == 1 for r in coop)
    # frozen BN (eval) and inference never take the fused training kernel
    seg.eval()
    R = PR.SegProgram(seg, (1, 3, 64, 64), training=False, with_grad=False, seg_size=(64, 64), dry_run=True)
    assert not any(getattr(r, "coop", False) for r in R.records if isinstance(r, PR.ConvBNRec))
