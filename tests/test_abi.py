"""CPU: the C-ABI shared library loads and exports every symbol include/sseg_b200.h declares, with the arity the ctypes
binding assumes; the binding fails loudly (no fallback) when the library is absent."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "sseg_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|long|void|const char\*)\s+(sseg_\w+)\s*\(([^;]*?)\)\s*;", hdr, flags=re.S):
        args = [a for a in m.group(2).split(",") if a.strip() and a.strip() != "void"]
        out[m.group(1)] = len(args)
    return out


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    from mit_semseg.engine import _C
    if not os.path.exists(_C.LIB_PATH):
        ge.build()
    L = _C.lib()
    decl = _declared()
    assert len(decl) >= 25
    for name, nargs in decl.items():
        assert hasattr(L, name), "libsseg_b200.so does not export %s" % name
        if name in _C._SIGNATURES:
            assert len(_C._SIGNATURES[name]) == nargs, "ctypes arity mismatch for %s" % name
    for name in _C._SIGNATURES:
        assert name in decl, "%s bound in _C.py but not declared in the header" % name
    assert L.sseg_version() >= 100


def test_no_silent_fallback_when_library_missing(monkeypatch):
    from mit_semseg.engine import _C
    monkeypatch.setattr(_C, "_lib", None)
    monkeypatch.setattr(_C, "LIB_PATH", "/nonexistent/libsseg_b200.so")
    with pytest.raises(_C.SsegError):
        _C.lib()


def test_cpu_tensors_are_rejected_not_emulated():
    import torch
    import torch.nn as nn
    from mit_semseg.models import ModelBuilder, SegmentationModule
    from mit_semseg.models import models as M, resnet as R
    enc = M.ResnetDilated(R.resnet18(pretrained=False), 8)
    dec = ModelBuilder.build_decoder("c1", fc_dim=512, num_class=150)
    seg = SegmentationModule(enc, dec, nn.NLLLoss(ignore_index=-1))
    feed = {"img_data": torch.zeros(1, 3, 64, 64), "seg_label": torch.zeros(1, 8, 8, dtype=torch.long)}
    with pytest.raises(RuntimeError, match="no CPU path"):
        seg(feed)


def test_single_process_multi_gpu_request_is_refused_not_degraded():
    """`python train.py --gpus 0-7` in ONE process (the reference's way, lib/nn/parallel/data_parallel.py:53-62) must not
    silently train on one eighth of the data: the wrapper raises and names the one-process-per-GPU launch line."""
    import torch
    from mit_semseg.lib.nn import UserScatteredDataParallel

    class Echo(torch.nn.Module):
        def forward(self, feed):
            return feed

    with pytest.raises(RuntimeError, match="torch.distributed.run"):
        UserScatteredDataParallel(Echo(), device_ids=[0, 1, 2, 3])
