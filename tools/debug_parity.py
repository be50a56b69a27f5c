#!/usr/bin/env python
"""Layer-by-layer parity of one engine training step against the CPU oracle (GPU box).
   python tools/debug_parity.py [--enc resnet50dilated] [--dec ppm_deepsup] [--fc 2048] [--n 2] [--hw 128]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "semantic-segmentation-pytorch_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
import torch.nn.functional as F  # noqa: E402


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--enc", default="resnet50dilated")
    ap.add_argument("--dec", default="ppm_deepsup")
    ap.add_argument("--fc", type=int, default=2048)
    ap.add_argument("--n", type=int, default=2)
    ap.add_argument("--hw", type=int, default=128)
    ap.add_argument("--emulate", default=None, help="bf16: oracle rounds at the engine's storage points")
    ap.add_argument("--gain", type=float, default=None, help="residual_gain of the synthetic weights")
    ap.add_argument("--brief", action="store_true")
    ap.add_argument("--bn-eval", action="store_true", help="BatchNorm layers in eval mode (fixed affine): removes the "
                    "batch-statistics coupling, so gradient wiring errors are not masked by small-batch BN chaos")
    args = ap.parse_args()
    from test_gpu_e2e import _build
    from mit_semseg.engine.program import SegProgram, ConvBNRec, StemRec, MaxPoolRec, ClassifierRec
    from oracle import segnet_oracle as O
    seg, esd, dsd, ds = _build(args.enc, args.dec, args.fc, residual_gain=args.gain)
    for m in seg.modules():
        if isinstance(m, nn.Dropout2d):
            m.p = 0.0
    seg.cuda().train()
    if args.bn_eval:
        for m in seg.modules():
            if isinstance(m, nn.modules.batchnorm._BatchNorm):
                m.eval()
    feed = O.synth_batch(args.n, args.hw, args.hw, 8, 1)
    prog = SegProgram(seg, tuple(feed["img_data"].shape), training=True, with_grad=True)
    prog.load_inputs(feed["img_data"].cuda(), feed["seg_label"].cuda())
    prog.run_eager()
    torch.cuda.synchronize()
    # ---- oracle with hooks on every conv/bn output: re-implement by monkeypatching O._cbr to record
    rec_o = {}
    orig_cbr = O._cbr

    def cbr(x, sd, conv, bn, st, stride=1, dilation=1, padding=0, relu=True):
        w = sd[conv + ".weight"] if x.shape[1] == 3 else O._qw(sd[conv + ".weight"], st)
        y = O._q(F.conv2d(x, w, sd.get(conv + ".bias"), stride, padding, dilation), st)
        z = O.batch_norm(y, sd, bn, st)
        rec_o[conv] = (y.detach(), z.detach())
        return O._q(F.relu(z), st) if relu else z

    O._cbr = cbr
    e = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in esd.items()}
    d = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in dsd.items()}
    l_ref, a_ref, feats, out = O.segmentation_forward(feed, e, d, args.enc, args.dec, O.BNState(not args.bn_eval, emulate=args.emulate), ds,
                                                      dropout_p=0.0, return_aux=True)
    l_ref.backward()
    O._cbr = orig_cbr
    print("loss %.5f vs %.5f ; acc %.4f vs %.4f" % (prog.out[0].item(), l_ref.item(), prog.out[1].item(), a_ref.item()))
    # map module -> name
    names = {}
    for prefix, net in (("", seg.encoder), ("", seg.decoder)):
        for n, m in net.named_modules():
            names[id(m)] = n
    print("config: bn_eval=%s emulate=%s gain=%s %s+%s n=%d hw=%d" % (args.bn_eval, args.emulate, args.gain, args.enc, args.dec, args.n, args.hw))
    print("%-28s %10s %10s" % ("conv (raw y) / applied", "rel(y)", "rel(a)"))
    for r in ([] if args.brief else prog.records):
        if isinstance(r, (ConvBNRec, StemRec)):
            nm = names[id(r.cw.mod)]
            if nm not in rec_o:
                continue
            y_o, z_o = rec_o[nm]
            y = r.y.float().cpu().permute(0, 3, 1, 2)
            ry = rel(y, y_o)
            ra = float("nan")
            if r.a is not None and getattr(r, "res", None) is None:
                a = r.a.t.float().cpu().permute(0, 3, 1, 2)
                zz = F.relu(z_o) if getattr(r, "relu", True) else z_o
                ra = rel(a, zz)
            print("%-28s %10.4f %10.4f   shape %s" % (nm, ry, ra, tuple(y.shape)))
    for i, (f, fo) in enumerate(zip(prog.feats, feats)):
        print("feat[%d] rel %.4f" % (i, rel(f.t.float().cpu().permute(0, 3, 1, 2), fo.detach())))
    heads = [r for r in prog.records if isinstance(r, ClassifierRec)]
    outs = out if isinstance(out, tuple) else (out,)
    for h, o in zip(heads, outs):
        lg = h.logits[..., :150].float().cpu().permute(0, 3, 1, 2)
        print("head log-prob rel %.4f ; logits std %.3f" % (rel(torch.log_softmax(lg, 1), o.detach()), lg.std().item()))
    # gradients
    grads = prog.param_grads()
    rows = []
    for prefix, net, sd in (("enc.", seg.encoder, e), ("dec.", seg.decoder, d)):
        for n, p in net.named_parameters():
            g = grads[p].float().cpu()
            gr = sd[n].grad
            rows.append((rel(g, gr), prefix + n, gr.norm().item()))
    rows.sort(reverse=True)
    print("worst gradients:")
    for r_, n, nr in rows[:(8 if args.brief else 25)]:
        print("  %-40s rel %.4f  |g| %.3e" % (n, r_, nr))
    import statistics
    print("median grad rel %.4f" % statistics.median(r[0] for r in rows))


if __name__ == "__main__":
    main()
