#!/usr/bin/env python
"""Phase-by-phase cycle breakdown of the tcgen05 implicit-GEMM kernel on the bench network's layer shapes.

Needs the development build of the library (cycle stamps compiled in):
    make -C semantic-segmentation-pytorch_b200/csrc trace
    SSEG_LIB=libsseg_b200_trace.so python tools/trace_igemm.py
Each CTA records clock64() at: kernel entry, prologue done, dependency wait done, first TMA issued, last TMA issued,
first full barrier seen by the MMA warp, last MMA committed, accumulator-ready seen by the epilogue, tile staged, statistics
done, stores done, exit. Printed: medians over the CTAs of one launch (cycles), after two warm launches of the same shape."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "semantic-segmentation-pytorch_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

SHAPES = [  # name, n, h, w, cin, cout, k, dil, stats
    ("layer3 conv1 1x1 1024->256", 2, 64, 64, 1024, 256, 1, 1, True),
    ("layer3 conv2 3x3d2 256->256", 2, 64, 64, 256, 256, 3, 2, True),
    ("layer3 conv3 1x1 256->1024", 2, 64, 64, 256, 1024, 1, 1, True),
    ("layer4 conv2 3x3d4 512->512", 2, 64, 64, 512, 512, 3, 4, True),
    ("layer4 conv3 1x1 512->2048", 2, 64, 64, 512, 2048, 1, 1, True),
    ("layer1 conv2 3x3 64->64", 2, 128, 128, 64, 64, 3, 1, True),
    ("conv_last 3x3 4096->512", 2, 64, 64, 4096, 512, 3, 1, True),
]
NAMES = ["entry", "prologue", "dep-wait", "1st TMA", "last TMA", "1st full", "last MMA", "acc seen", "staged", "bar", "stats",
         "stores issued", "epilogue end", "sync", "dealloc"]   # (stamp 11 = after the tile stores, which now precede the sums)


def main():
    from mit_semseg.engine import _C, ops
    L = _C.lib()
    assert hasattr(L, "sseg_debug_read_trace"), "run with SSEG_LIB=libsseg_b200_trace.so (make -C .../csrc trace)"
    L.sseg_debug_read_trace.argtypes = [ctypes.c_void_p, ctypes.c_long]
    g = torch.Generator(device="cuda").manual_seed(0)
    for name, n, h, w, cin, cout, k, d, stats in SHAPES:
        x = torch.randn(n, h, w, cin, device="cuda", generator=g).bfloat16()
        K = k * k * cin
        wt = (torch.randn(cout, K, device="cuda", generator=g) * (2.0 / K) ** 0.5).bfloat16()
        out = torch.empty(n, h, w, cout, device="cuda", dtype=torch.bfloat16)
        ssum, ssq = torch.zeros(cout, device="cuda"), torch.zeros(cout, device="cuda")
        geom = ops.make_geom([x], ops.conv_taps(k, d))
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for it in range(3):
            if it == 2:
                torch.cuda.synchronize()
                L.sseg_debug_clear_trace()
                ev0.record()
            ops.conv_igemm(geom, wt, cout, out, stat_sum=ssum, stat_sqsum=ssq)
        ev1.record()
        torch.cuda.synchronize()
        buf = (ctypes.c_longlong * (4096 * 16))()
        _C.check(L.sseg_debug_read_trace(buf, 4096 * 16))
        t = torch.tensor(list(buf), dtype=torch.int64).view(4096, 16)
        live = t[t[:, 0] > 0]
        rel = (live - live[:, :1]).double()
        med = rel.median(0).values
        flops = 2.0 * n * h * w * cout * K
        print("%s: %d CTAs traced, launch %.2f us (event), %.0f TFLOP/s" % (name, live.shape[0], ev0.elapsed_time(ev1) * 1e3,
                                                                            flops / (ev0.elapsed_time(ev1) * 1e-3) / 1e12))
        print("   " + "  ".join("%s %d" % (NAMES[i], int(med[i])) for i in range(1, 15) if med[i] > 0))
        span = (live[:, 13].max() - live[:, 0].min()).item()
        print("   CTA lifetime median %d cycles; first entry -> last exit %d cycles (per-SM clocks are not synchronised: indicative)"
              % (int(med[13]), span))


if __name__ == "__main__":
    main()
