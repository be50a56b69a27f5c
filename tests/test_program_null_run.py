"""CPU: every closure of a schedule is EXECUTED with the kernels replaced by argument-checking no-ops.

`libsseg_b200.so` is swapped for a stand-in whose entry points validate each call against the ctypes signatures of the
real library (engine/_C.py::_SIGNATURES: argument count and convertibility) and return 0; the torch bookkeeping ops of
the schedule (copies, memsets, scaling) run for real on CPU tensors. Nothing is computed - this is not an execution
path - but every Python-level mistake in a closure (a wrong variable, a wrapper called with the wrong arguments, a
non-dense view handed to a dense-only kernel, a stale attribute) fails here, without a GPU, for the default schedule
and for every opt-in switch."""
import ctypes
import itertools

import pytest
import torch
import torch.nn as nn

from test_program_dry import _seg


class _NullLib:
    """Stand-in for ctypes.CDLL(libsseg_b200.so): checks marshalling, launches nothing."""

    def __init__(self, signatures):
        self.sig = signatures
        self.calls = {}

    def __getattr__(self, name):
        if name == "sseg_launch_count":
            return lambda: sum(self.calls.values())
        if name == "sseg_launch_count_reset":
            return lambda: self.calls.clear()
        if name == "sseg_last_error":
            return lambda: b"null library"
        if name not in self.sig:
            raise AttributeError(name)
        argtypes = self.sig[name]

        def call(*args):
            assert len(args) == len(argtypes), "%s: %d arguments for %d parameters" % (name, len(args), len(argtypes))
            for i, (a, t) in enumerate(zip(args, argtypes)):
                try:
                    t.from_param(a)
                except (TypeError, ctypes.ArgumentError) as exc:
                    raise AssertionError("%s: argument %d (%r) is not a %s: %s" % (name, i, a, t.__name__, exc))
            self.calls[name] = self.calls.get(name, 0) + 1
            return 1 if name.endswith("_fits") else 0
        return call


@pytest.fixture
def null_lib(monkeypatch):
    from mit_semseg.engine import _C, ops
    lib = _NullLib(_C._SIGNATURES)
    monkeypatch.setattr(_C, "lib", lambda: lib)
    monkeypatch.setattr(ops, "_stream", lambda: None)
    for name in ("SSEG_BRANCH_STREAMS", "SSEG_FOLD_BN_EVAL", "SSEG_OVERLAP_RELAYOUT", "SSEG_COOP_BN"):
        monkeypatch.delenv(name, raising=False)
    return lib


def _run(prog):
    prog.dry_run, prog.serial = False, True   # serial: side / branch streams run inline (no CUDA streams on this box)
    prog.run_eager()
    prog.run_eager()


TRAIN_NETS = [("resnet50dilated", "ppm_deepsup", 2048, 8), ("resnet18dilated", "c1_deepsup", 512, 8),
              ("resnet50", "upernet", 2048, 4), ("hrnetv2", "c1", 720, 4)]
SWITCHES = ["SSEG_BRANCH_STREAMS", "SSEG_OVERLAP_RELAYOUT", "SSEG_COOP_BN"]


@pytest.mark.parametrize("enc,dec,fc,stride", TRAIN_NETS)
@pytest.mark.parametrize("switches", [()] + [(s,) for s in SWITCHES] + [tuple(SWITCHES)])
def test_training_schedule_executes(enc, dec, fc, stride, switches, null_lib, monkeypatch):
    from mit_semseg.engine import program as PR
    for s in switches:
        monkeypatch.setenv(s, "1")
    seg = _seg(enc, dec, fc)
    seg.train()
    P = PR.SegProgram(seg, (2, 3, 64, 64), training=True, with_grad=True, dry_run=True)
    P.load_inputs(torch.randn(2, 3, 64, 64), torch.randint(-1, 150, (2, 64 // stride, 64 // stride)))
    _run(P)
    calls = null_lib.calls
    nconv = sum(isinstance(m, nn.Conv2d) for m in seg.modules())
    assert calls["sseg_conv_wgrad"] == 2 * (nconv - 1) and calls["sseg_stem_conv_wgrad"] == 2   # two runs
    fwd_convs = sum(calls.get(k, 0) for k in ("sseg_conv_igemm", "sseg_conv_bn_train", "sseg_conv_igemm_bnbwd",
                                              "sseg_conv_igemm_bnbwd_res", "sseg_conv_igemm_bnfin", "sseg_conv_dgrad_bn"))
    assert fwd_convs >= 2 * 2 * (nconv - 1) - 8          # forward + data gradient of every conv but the stem (first layers have no dgrad)
    if "SSEG_COOP_BN" in switches:
        assert calls["sseg_conv_bn_train"] > 0 and calls["sseg_conv_dgrad_bn"] > 0
        assert calls.get("sseg_bn_bwd_apply", 0) < 2 * sum(isinstance(m, nn.modules.batchnorm._BatchNorm) for m in seg.modules())
    grads = P.param_grads()
    assert all(grads[p].shape == p.shape for p in seg.parameters())
    # frozen BN (the reference's fix_bn) takes the eval branches of the same schedule
    for m in seg.modules():
        if isinstance(m, nn.modules.batchnorm._BatchNorm):
            m.eval()
    _run(PR.SegProgram(seg, (2, 3, 64, 64), training=True, with_grad=True, dry_run=True))


@pytest.mark.parametrize("enc,dec,fc", [("resnet18dilated", "ppm_deepsup", 512), ("resnet50", "upernet", 2048), ("hrnetv2", "c1", 720),
                                        ("mobilenetv2dilated", "c1_deepsup", 320)])
@pytest.mark.parametrize("fold", ["0", "1"])
def test_inference_schedules_execute(enc, dec, fc, fold, null_lib, monkeypatch):
    from mit_semseg.engine import program as PR
    monkeypatch.setenv("SSEG_FOLD_BN_EVAL", fold)
    seg = _seg(enc, dec, fc)
    seg.eval()
    P = PR.SegProgram(seg, (1, 3, 64, 96), training=False, with_grad=False, seg_size=(64, 96), dry_run=True)
    _run(P)
    assert ("sseg_conv_igemm_affine" in null_lib.calls) == (fold == "1") or enc == "mobilenetv2dilated"   # always folded
    if fold == "1":
        assert null_lib.calls.get("sseg_bn_apply", 0) <= 2      # only the stem's BN (its conv is not a tcgen05 GEMM)
    # module-level encoder / decoder programs
    E = PR.SegProgram(None, (1, 3, 64, 96), training=False, with_grad=False, part="encoder", enc=seg.encoder, dry_run=True)
    _run(E)
    shapes = tuple(tuple(t.shape) for t in E.feat_out)
    D = PR.SegProgram(None, None, training=False, with_grad=False, part="decoder", dec=seg.decoder, feat_shapes=shapes,
                      dry_run=True)
    D.load_features([torch.randn(s) for s in shapes])
    _run(D)
    # multi-scale head: two scales accumulate into one score map
    scores = torch.zeros(1, 150, 64, 96)
    for hw in ((64, 96), (96, 128)):
        _run(PR.SegProgram(seg, (1, 3) + hw, training=False, with_grad=False, seg_size=(64, 96), dry_run=True, head_out=scores,
                           head_weight=0.5))


@pytest.fixture
def null_engine(null_lib, monkeypatch):
    """The public entry points (engine/functional.py) on CPU tensors: programs are built as dry-run schedules and executed
    against the null library, CUDA-graph capture is skipped, the `is_cuda` guards are told the tensors live on a GPU."""
    from mit_semseg.engine import functional as EF
    from mit_semseg.engine import program as PR
    real = PR.SegProgram

    def factory(*a, **k):
        k["dry_run"] = True
        prog = real(*a, **k)
        prog.dry_run, prog.serial = False, True
        return prog
    monkeypatch.setattr(EF, "SegProgram", factory)
    monkeypatch.setattr(real, "capture", lambda self, warm=True: None)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))
    from mit_semseg.engine import accurate as ACC
    real_acc = ACC.AccurateInference

    def acc_factory(*a, **k):
        prog = real_acc(*a, dry_run=True, **k)
        prog.dry_run = False
        return prog
    monkeypatch.setattr(ACC, "AccurateInference", acc_factory)
    monkeypatch.setattr(real_acc, "capture", lambda self: None)
    return null_lib


def test_public_entry_points_execute(null_engine):
    """SegmentationModule(feed) -> (loss, acc) -> loss.backward() hands every parameter a gradient of its own shape;
    segSize=... returns probabilities; module-level encoder / decoder calls and the fused multi-scale loop run."""
    from mit_semseg.engine import functional as EF
    seg = _seg("resnet18dilated", "ppm_deepsup", 512)
    seg.train()
    feed = {"img_data": torch.randn(2, 3, 64, 64), "seg_label": torch.randint(-1, 150, (2, 8, 8))}
    loss, acc = seg(feed)
    assert loss.dim() == 0 and acc.dim() == 0 and loss.requires_grad and not acc.requires_grad
    (loss * 2).backward()
    assert all(p.grad is not None and p.grad.shape == p.shape for p in seg.parameters())
    loss2, _ = seg(feed)                      # second call: same cached program
    assert len(seg.__dict__["_b200_programs"]) == 1
    with torch.no_grad():
        l3, _ = seg(feed)                     # no-grad training-mode call builds a forward-only program
        assert not l3.requires_grad and len(seg.__dict__["_b200_programs"]) == 2
    seg.eval()
    seg.decoder.use_softmax = True
    with torch.no_grad():
        probs = seg({"img_data": torch.randn(1, 3, 64, 96)}, segSize=(64, 96))
        assert probs.shape == (1, 150, 64, 96)
        feats = seg.encoder(torch.randn(1, 3, 64, 96), return_feature_maps=True)
        assert [tuple(f.shape) for f in feats] == [(1, 64, 16, 24), (1, 128, 8, 12), (1, 256, 8, 12), (1, 512, 8, 12)]
        out = seg.decoder(feats, segSize=(64, 96))
        assert out.shape == (1, 150, 64, 96)
        imgs = [torch.randn(1, 3, 64, 96), torch.randn(1, 3, 96, 128), torch.randn(1, 3, 128, 160)]
        for _ in range(3):
            scores = EF.multiscale_inference(seg, imgs, (64, 96))
        assert scores.shape == (1, 150, 64, 96)
        # fp32-accurate mode through the same entry point
        import os
        os.environ["SSEG_ACCURATE_INFERENCE"] = "1"
        try:
            for _ in range(3):
                probs = seg({"img_data": torch.randn(1, 3, 64, 96)}, segSize=(64, 96))
        finally:
            os.environ.pop("SSEG_ACCURATE_INFERENCE")
        assert probs.shape == (1, 150, 64, 96) and null_engine.calls["sseg_split_affine"] > 0
    seg.decoder.use_softmax = False
    with pytest.raises(RuntimeError, match="use_softmax"):
        seg({"img_data": torch.randn(1, 3, 64, 96)}, segSize=(64, 96))

