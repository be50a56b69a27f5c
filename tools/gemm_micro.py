#!/usr/bin/env python
"""Micro-benchmark of the conv GEMM kernels on the bench network's layer shapes: 30 back-to-back launches of one layer
(programmatic dependent launch as in the step), CUDA events around the batch -> us per launch and TFLOP/s. The tile / kernel
variant is chosen by the library's environment knobs, so variants are compared by running this script under different
settings:
    python tools/gemm_micro.py                       # defaults
    SSEG_IGEMM_2CTA=2 python tools/gemm_micro.py     # CTA pairs (tcgen05 cta_group::2) wherever the M tiles pair up
    SSEG_IGEMM_N256=0 python tools/gemm_micro.py     # 128-wide tiles only
Shapes: forward convs (with BN statistics in the epilogue), the data gradients of the same layers and weight gradients."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "semantic-segmentation-pytorch_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

SHAPES = [  # name, n, h, w, cin, cout, k, dil
    ("layer1 conv2 3x3 64->64", 2, 128, 128, 64, 64, 3, 1),
    ("layer2 conv3 1x1 128->512", 2, 64, 64, 128, 512, 1, 1),
    ("layer3 conv1 1x1 1024->256", 2, 64, 64, 1024, 256, 1, 1),
    ("layer3 conv2 3x3d2 256->256", 2, 64, 64, 256, 256, 3, 2),
    ("layer3 conv3 1x1 256->1024", 2, 64, 64, 256, 1024, 1, 1),
    ("layer4 conv1 1x1 2048->512", 2, 64, 64, 2048, 512, 1, 1),
    ("layer4 conv2 3x3d4 512->512", 2, 64, 64, 512, 512, 3, 4),
    ("layer4 conv3 1x1 512->2048", 2, 64, 64, 512, 2048, 1, 1),
    ("deepsup 3x3 1024->512", 2, 64, 64, 1024, 512, 3, 1),
    ("conv_last 3x3 4096->512", 2, 64, 64, 4096, 512, 3, 1),
]
REPS = 30


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPS):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / REPS


def main():
    from mit_semseg.engine import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    knobs = {k: v for k, v in os.environ.items() if k.startswith("SSEG_")}
    print("knobs:", knobs or "(defaults)")
    print("%-30s %10s %8s %10s %8s %10s %8s" % ("layer", "fwd us", "TF/s", "dgrad us", "TF/s", "wgrad us", "TF/s"))
    tot = [0.0, 0.0, 0.0]
    for name, n, h, w, cin, cout, k, d in SHAPES:
        T = k * k
        x = torch.randn(n, h, w, cin, device="cuda", generator=g).bfloat16()
        wf = (torch.randn(cout, T * cin, device="cuda", generator=g) * (2.0 / (T * cin)) ** 0.5).bfloat16()
        y = torch.empty(n, h, w, cout, device="cuda", dtype=torch.bfloat16)
        ssum, ssq = torch.zeros(cout, device="cuda"), torch.zeros(cout, device="cuda")
        geom = ops.make_geom([x], ops.conv_taps(k, d))
        t_f = timed(lambda: ops.conv_igemm(geom, wf, cout, y, stat_sum=ssum, stat_sqsum=ssq))
        # data gradient: dy [.., cout] x wd [cin][T * cout] -> dx [.., cin], taps mirrored
        dy = torch.randn(n, h, w, cout, device="cuda", generator=g).bfloat16()
        wd = (torch.randn(cin, T * cout, device="cuda", generator=g) * 0.02).bfloat16()
        dx = torch.empty(n, h, w, cin, device="cuda", dtype=torch.bfloat16)
        dh, dw = ops.conv_taps(k, d)
        gd = ops.make_geom([dy], ([-v for v in dh], [-v for v in dw]), tap_koff=[t * cout for t in range(T)])
        t_d = timed(lambda: ops.conv_igemm(gd, wd, cin, dx, n_store=cin))
        gw = torch.zeros(cout, T * cin, device="cuda", dtype=torch.float32)
        t_w = timed(lambda: ops.conv_wgrad(geom, dy, cout, gw))
        fl = 2.0 * n * h * w * cout * T * cin
        print("%-30s %10.2f %8.0f %10.2f %8.0f %10.2f %8.0f" % (name, t_f, fl / t_f / 1e6, t_d, fl / t_d / 1e6, t_w, fl / t_w / 1e6))
        tot[0] += t_f
        tot[1] += t_d
        tot[2] += t_w
    print("%-30s %10.2f %8s %10.2f %8s %10.2f" % ("sum", tot[0], "", tot[1], "", tot[2]))


if __name__ == "__main__":
    main()
