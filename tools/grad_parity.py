#!/usr/bin/env python
"""Per-parameter gradient parity of one train-mode step against the oracle (GPU box, or the ABI emulator).

    python tools/grad_parity.py [--enc resnet18dilated] [--dec ppm_deepsup] [--fc 512] [--hw 128] [--emulate bf16|none]
                                [--calibrate] [--bias-shift 0.0] [--repeat 3]

Prints, per parameter, |g_engine| / |g_oracle| and the cosine, sorted by the parameter's share of the squared-norm
excess; `--repeat` re-runs the engine step to show the run-to-run noise of the atomics order."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "semantic-segmentation-pytorch_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--enc", default="resnet18dilated")
    ap.add_argument("--dec", default="ppm_deepsup")
    ap.add_argument("--fc", type=int, default=512)
    ap.add_argument("--n", type=int, default=2)
    ap.add_argument("--hw", type=int, default=128)
    ap.add_argument("--emulate", default="bf16")
    ap.add_argument("--gain", type=float, default=0.25)
    ap.add_argument("--calibrate", action="store_true")
    ap.add_argument("--bias-shift", type=float, default=0.0)
    ap.add_argument("--repeat", type=int, default=1)
    ap.add_argument("--top", type=int, default=25)
    args = ap.parse_args()
    from test_gpu_e2e import _build
    from mit_semseg.engine.program import SegProgram
    from oracle import segnet_oracle as O
    feed = O.synth_batch(args.n, args.hw, args.hw, 8, 7)
    seg, esd, dsd, ds = _build(args.enc, args.dec, args.fc, residual_gain=args.gain, bias_shift=args.bias_shift,
                               calibrate_on=feed if args.calibrate else None)
    for m in seg.modules():
        if isinstance(m, nn.Dropout2d):
            m.p = 0.0
    seg.cuda().train()
    prog = SegProgram(seg, tuple(feed["img_data"].shape), training=True, with_grad=True)
    e = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in esd.items()}
    d = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in dsd.items()}
    st = O.BNState(True, emulate=None if args.emulate == "none" else args.emulate)
    l_ref, a_ref = O.segmentation_forward(feed, e, d, args.enc, args.dec, st, ds, dropout_p=0.0)
    l_ref.backward()
    names = [("enc." + k, p, e[k].grad) for k, p in seg.encoder.named_parameters()] + \
            [("dec." + k, p, d[k].grad) for k, p in seg.decoder.named_parameters()]
    prev = None
    for it in range(args.repeat):
        prog.load_inputs(feed["img_data"].cuda(), feed["seg_label"].cuda())
        prog.run_eager()
        torch.cuda.synchronize()
        grads = prog.param_grads()
        rows, dot, n1, n2 = [], 0.0, 0.0, 0.0
        cur = {}
        for name, p, gr in names:
            g = grads[p].float().cpu().double().flatten()
            gr = gr.double().flatten()
            cur[name] = g
            a, b, c = torch.dot(g, g).item(), torch.dot(gr, gr).item(), torch.dot(g, gr).item()
            dot, n1, n2 = dot + c, n1 + a, n2 + b
            rows.append((a - b, name, (a / (b + 1e-300)) ** 0.5, c / ((a * b) ** 0.5 + 1e-300), b ** 0.5))
        print("run %d: loss %.5f (oracle %.5f)  grad cosine %.4f  norm ratio %.4f" %
              (it, prog.out[0].item(), l_ref.item(), dot / (n1 * n2) ** 0.5, (n1 / n2) ** 0.5))
        if prev is not None:
            dd = sum(((cur[k] - prev[k]) ** 2).sum().item() for k in cur) ** 0.5
            print("        run-to-run |dg| / |g| = %.3e" % (dd / n1 ** 0.5))
        prev = cur
        if it == 0:
            print("%-46s %9s %8s %10s %8s" % ("parameter", "|g|/|ref|", "cos", "|ref|", "excess%"))
            for ex, name, ratio, cos, nb in sorted(rows, key=lambda r: -abs(r[0]))[:args.top]:
                print("%-46s %9.4f %8.4f %10.3e %7.2f%%" % (name, ratio, cos, nb, 100 * ex / n2))


if __name__ == "__main__":
    main()
