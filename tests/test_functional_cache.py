"""Host-side bookkeeping of the public entry points (engine/functional.py): the cached flat module / parameter lists that
replace nn.Module's tree walks on the per-step path, and SegmentationModule.zero_grad on top of them. Pure torch, no library."""
import torch
import torch.nn as nn

from test_program_dry import _seg


def test_cached_tree_follows_the_module_tree(monkeypatch):
    from mit_semseg.engine import functional as EF
    from mit_semseg.models import ModelBuilder
    seg = _seg("resnet18dilated", "ppm_deepsup", 512)
    modules, params = EF._tree(seg)
    assert [id(m) for m in modules] == [id(m) for m in seg.modules()]
    assert [id(p) for p in params] == [id(p) for p in seg.parameters()]
    assert EF._tree(seg)[1] is params                      # served from the cache
    seg.__dict__["_b200_programs"] = {"stale": object()}
    # a direct child is replaced: noticed at once, the programs compiled from the old tree are dropped
    seg.decoder = ModelBuilder.build_decoder("c1_deepsup", fc_dim=512, num_class=150)
    modules2, params2 = EF._tree(seg)
    assert [id(p) for p in params2] == [id(p) for p in seg.parameters()] and len(params2) != len(params)
    assert "_b200_programs" not in seg.__dict__
    # a change deep inside the tree: picked up by the periodic full re-walk
    monkeypatch.setattr(EF, "_TREE_RECHECK", 3)
    seg.__dict__.pop("_b200_tree")
    EF._tree(seg)
    seg.__dict__["_b200_programs"] = {"stale": object()}
    seg.encoder.layer1[0].conv1 = nn.Conv2d(64, 64, 3, padding=1, bias=True)     # adds a bias parameter
    for _ in range(4):
        got = EF._tree(seg)[1]
    assert [id(p) for p in got] == [id(p) for p in seg.parameters()]
    assert "_b200_programs" not in seg.__dict__
    # training flags are read through the cached module list
    f0 = EF._flags(seg)
    seg.encoder.layer2.eval()
    assert EF._flags(seg) != f0


def test_zero_grad_semantics():
    seg = _seg("resnet18dilated", "c1", 512)
    ref = nn.Module.zero_grad
    for set_to_none in (True, False):
        for p in seg.parameters():
            p.grad = torch.ones_like(p)
        frozen = next(seg.parameters())
        frozen.grad = None
        seg.zero_grad(set_to_none=set_to_none)
        for p in seg.parameters():
            if set_to_none or p is frozen:
                assert p.grad is None
            else:
                assert p.grad is not None and not p.grad.requires_grad and float(p.grad.abs().sum()) == 0.0
    # same observable result as the generic implementation
    for p in seg.parameters():
        p.grad = torch.ones_like(p)
    ref(seg)
    assert all(p.grad is None for p in seg.parameters())
