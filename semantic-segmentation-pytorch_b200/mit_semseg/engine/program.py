"""Step programs: a SegmentationModule + an input shape compiled into a fixed schedule of sm_100a kernel launches.

A program is built ONCE per (module, input shape, mode): it walks the module tree (reading every conv's stride /
dilation / padding from the nn.Conv2d that owns the weights, so ResnetDilated's rewrites are honoured), allocates all
activations, statistics and gradient buffers up front, and records two lists of closures:

    fwd : weight re-layout -> stem conv -> (conv [tcgen05 implicit GEMM, BN statistics in the epilogue]
          -> [NCCL all-reduce of the statistics when synchronised] -> BN finalize -> fused BN/residual/ReLU/dropout apply)*
          -> PPM cascade (adaptive pools, 1x1 convs, bilinear up-sampling, virtual concat) -> classifier(s)
          -> fused log-softmax / NLL / pixel-accuracy
    bwd : the exact adjoint, in reverse: loss gradient -> (BN backward reduce -> [all-reduce] -> BN backward apply
          -> weight gradient GEMM (K = pixels, split-K) -> data gradient implicit GEMM)* -> gradient bucket all-reduce

Running the program is just calling the closures in order on the current stream: no shapes are inspected, no memory
is allocated, nothing synchronises — so a whole training step can be captured into ONE CUDA graph
(`SegProgram.capture()`), which is how `bench.py` and `segmentation_train_step` run it.

Activations are NHWC bf16, accumulation is fp32 (TMEM), BN statistics / loss / all parameter gradients are fp32.
Reference call sites are cited on each record class.
"""
import math
import os

import torch
import torch.nn as nn
from torch.nn.modules.batchnorm import _BatchNorm

from . import ops


def _pad(x, m):
    return (x + m - 1) // m * m


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


def _coop_fits_estimate(n, h, w, k, channels):
    """CPU-side stand-in for sseg_conv_bn_train_fits / sseg_conv_dgrad_bn_fits (dry-run schedules only): tiles of 128 pixels
    x 64|128 channels, one persistent CTA per SM (148; SSEG_DRY_RUN_SMS overrides it for schedules built for the CPU
    simulator's smaller "device"), at most 512 tensor-memory columns per CTA."""
    sms = int(os.environ.get("SSEG_DRY_RUN_SMS", "148"))
    if k == 1:
        m_tiles = math.ceil(n * h * w / 128)
    else:
        bw = min(128, 1 << max(3, (w - 1).bit_length()))
        m_tiles = n * math.ceil(h / (128 // bw)) * math.ceil(w / bw)
    cp = _pad(channels, 8)
    cols = 64 if (channels <= 64 or m_tiles * math.ceil(cp / 128) <= 80) else 128
    return math.ceil(m_tiles * math.ceil(cp / cols) / sms) * cols <= 512


class Act:
    """An activation tensor (NHWC bf16) plus, during backward construction, its gradient buffer."""
    __slots__ = ("t", "tp", "g", "gw", "uses", "producer", "gcount")

    def __init__(self, t, tp=None):
        # t: the logical tensor (c channels; what convolutions read). tp: the same storage with the channel count
        # rounded up to 8 (zero pad channels) — what the 8-channel-vectorised elementwise kernels and gradient
        # writers see. They differ only for widths like HRNet+C1's 180-channel hidden layer.
        self.t, self.tp, self.g, self.gw = t, (t if tp is None else tp), None, False
        self.uses = 0          # number of consumers in the forward schedule
        self.gcount = 0        # gradient contributions scheduled so far (backward construction): complete at == uses
        self.producer = None   # the ConvBNRec / StemRec that wrote it (if any)


class ConvW:
    """Per-nn.Conv2d engine state: bf16 GEMM operands (forward / data-gradient layouts) and the fp32 gradient slot."""

    def __init__(self, mod):
        self.mod = mod
        self.O, self.I, kh, kw = mod.weight.shape
        assert kh == kw and mod.groups == 1, "square, ungrouped convolutions only"
        self.k, self.T = kh, kh * kw
        self.stride, self.dil, self.pad = mod.stride[0], mod.dilation[0], mod.padding[0]
        assert mod.stride[0] == mod.stride[1] and self.pad == self.dil * (self.k // 2), "only 'same' padding is supported"
        self.Opad = _pad(self.O, 64)
        self.ldf = _pad(self.T * self.I, 8)   # forward operand row pitch (TMA wants 16-byte multiples)
        self.wf = self.wd = self.gw = self.gb = None  # views, assigned by SegProgram._alloc_params
        # Every convolution the GEMM kernels run keeps its fp32 MASTER weight channels-last in memory ([O][kh][kw][I]; the
        # logical shape, state-dict values and checkpoints are unchanged - torch.channels_last is an ordinary memory
        # format): that is the layout the weight-gradient GEMM writes ([O][tap * I]), so `weight.grad` is a VIEW of the
        # flat gradient buffer and the per-step gradient re-layout pass (0.2 ms of a 6.4 ms step) does not exist; the
        # forward operand becomes a plain fp32 -> bf16 cast. (The 3-channel stem conv is read by its own kernel as OIHW.)
        self.cl = self.I != 3
        if self.cl and self.T > 1 and not mod.weight.data.is_contiguous(memory_format=torch.channels_last):
            mod.weight.data = mod.weight.data.contiguous(memory_format=torch.channels_last)


class BNS:
    """Per-BatchNorm engine state: statistics / normalisation vectors and gradient slots (dgamma, dbeta)."""

    def __init__(self, mod):
        self.mod, self.C = mod, mod.num_features
        self.Cp = _pad(self.C, 8)  # vector length seen by the 8-channel-vectorised kernels (pad entries stay 0)
        self.stats = self.mean = self.invstd = self.scale = self.shift = self.dgamma = self.dbeta = None


class SegProgram:
    def __init__(self, seg, img_shape, training, with_grad=True, seg_size=None, dropout_masks=None, part="full",
                 enc=None, dec=None, feat_shapes=None, dry_run=False, head_out=None, head_weight=1.0):
        """seg: SegmentationModule.  img_shape: (N, 3, H, W).  training: module.training (BN/dropout behaviour).
        with_grad: also build the backward schedule.  seg_size: inference branch (probabilities at seg_size).
        dropout_masks: optional {'main': [N,512] 0/1, 'deepsup': ...} to inject the Dropout2d draws (tests).
        part: "full" (encoder + decoder + loss/head) | "encoder" (module called on its own: fp32 NCHW feature maps out)
              | "decoder" (fp32 NCHW feature maps of `feat_shapes` in, log-probs / probabilities out); the partial
              programs are forward-only.
        head_out / head_weight: inference only — accumulate head_weight * probabilities into this fp32 [N,C,*seg_size]
              tensor instead of writing a fresh one (`scores += pred / len(scales)`, eval.py:72).
        dry_run: build the schedule only (shape inference, buffer plan, record wiring) — used by the CPU tests of the
              host logic; such a program refuses to run."""
        self.dry_run = bool(dry_run)
        self.head_out, self.head_weight = head_out, float(head_weight)
        self.seg = seg
        self.part = part
        self.enc = enc if enc is not None else (seg.encoder if seg is not None else None)
        self.dec = dec if dec is not None else (seg.decoder if seg is not None else None)
        if part == "encoder":
            self.dec = None
        elif part == "decoder":
            self.enc = None
        if part != "full":
            with_grad = False
        self.feat_shapes = feat_shapes
        if part == "decoder":
            img_shape = (feat_shapes[0][0], 3, 0, 0)
            # use only the maps the decoder reads (the last one, and the one before it for deep supervision)
        self.N, _, self.H, self.W = img_shape
        self.training = bool(training)
        self.inference = seg_size is not None
        self.seg_size = seg_size
        self.with_grad = bool(with_grad) and not self.inference
        self.dev = next((self.enc if self.enc is not None else self.dec).parameters()).device
        assert self.dev.type == "cuda" or self.dry_run, "the B200 engine runs on CUDA devices only (no CPU fallback)"
        self.injected_masks = dropout_masks
        self.dist = _dist()
        self.world = self.dist.get_world_size() if self.dist else 1
        self.fwd, self.bwd, self.records = [], [], []
        self.pool_groups = {}
        import os as _os
        self.fuse_bnbwd = _os.environ.get("SSEG_FUSE_BNBWD", "1") != "0"
        self.fuse_bnbwd_res = _os.environ.get("SSEG_FUSE_BNBWD_RES", "1") != "0"   # ... also for residual-block outputs
        # multi-GPU, opt-in: the SyncBN backward exchange (flag handshake + pooling of the partial sums out of peer memory)
        # inside the BN-backward apply kernel instead of a kernel of its own. Correct (dist_check, two-process tests) but
        # SLOWER on 2 x B200: 7.50 ms/step (one wave of blocks) / 8.19 (four waves) against 6.91 with the one-block exchange
        # kernel - every block of a full-GPU kernel then sits on the handshake + NVLink round trip, and the spinning blocks
        # take the SM slots the side-stream weight-gradient GEMMs would use (profiles/r2_summary.md section 6)
        # SyncBN exchange protocol between the GPUs: push ("LL": 8-byte {value, step} messages into the peers' inboxes, one
        # NVLink one-way latency) or, SSEG_PEER_LL=0, flag handshake + loads out of peer memory
        self.peer_ll = _os.environ.get("SSEG_PEER_LL", "1") != "0"
        self.defer_running = _os.environ.get("SSEG_DEFER_RUNNING", "1") != "0"
        self.deferred_running = []
        self.peer_fuse_bwd = _os.environ.get("SSEG_PEER_FUSE_BWD", "0") != "0"
        # measured on B200: fusing finalize into apply does NOT pay (6.92 vs 6.79 ms/step): with programmatic dependent
        # launch the tiny finalize kernel already overlaps the conv's tail, while the fused prologue delays every
        # block's streaming phase. Kept as an opt-in.
        self.fuse_finalize = _os.environ.get("SSEG_FUSE_FINALIZE", "0") == "1"
        self.keep = []  # anything that must stay alive (geometry structs hold raw pointers)
        self.graph = None
        # weight-gradient GEMMs are off the critical path of the backward pass (nothing downstream reads them until
        # the gradient bucket): they run on a side stream and fill the SMs the data-gradient chain leaves idle
        self.side = torch.cuda.Stream(self.dev) if self.dev.type == "cuda" else None
        # Independent sub-chains of small kernels (the four PPM pyramid branches; the parallel branches of an HRNet
        # module) can run on streams of their own inside the step graph instead of back to back: each chain is a
        # handful of latency-bound launches (8..200 CTAs) that leave most of the 148 SMs idle.
        # Measured on B200 (profiles/r2_first_run.log): 6.51 -> 6.21 ms/step (R50+PPM), 15.8 -> 10.4 ms (HRNetV2+C1): default on;
        # SSEG_BRANCH_STREAMS=0 turns it off.
        self.use_branches = _os.environ.get("SSEG_BRANCH_STREAMS", "1") == "1"
        # Inference programs: BatchNorm with running statistics is a per-channel affine, so conv -> BN -> (+shortcut) ->
        # ReLU runs as ONE kernel (sseg_conv_igemm_affine) and the raw conv output never exists. Opt-in until measured.
        self.fold_bn_eval = _os.environ.get("SSEG_FOLD_BN_EVAL", "0") == "1"
        # Weight re-layout off the critical path (opt-in until measured): the forward operands of everything but the first
        # layers and ALL data-gradient operands are produced on the side stream while the stem / first stages run, and
        # the gradient re-layout of a bucket follows its weight-gradient GEMMs on the side stream instead of the end.
        self.overlap_relayout = _os.environ.get("SSEG_OVERLAP_RELAYOUT", "0") == "1" and part == "full"
        self.split_prep = _os.environ.get("SSEG_SPLIT_PREP", "0") != "0" and not self.overlap_relayout
        # conv + train-mode BN (+shortcut, ReLU, dropout) as ONE persistent kernel with an in-kernel grid barrier
        # (sseg_conv_bn_train) for every layer whose tiles fit the SMs' tensor memory; single-GPU F.batch_norm branch only.
        # Opt-in until it has run on B200.
        self.coop_bn = _os.environ.get("SSEG_COOP_BN", "0") == "1"
        # BatchNorm finalize by the last CTA of the conv launch (single-GPU train mode). Measured on B200
        # (profiles/r2_summary.md): 6.33 ms/step against 5.94 with the separate bn_finalize launch - every CTA pays a
        # __threadfence() + ticket round trip behind its statistics atomics, while the tiny finalize kernel's launch already
        # overlaps the conv's tail under programmatic dependent launch. Opt-in: SSEG_FUSE_BNFIN=1.
        self.fuse_bnfin = _os.environ.get("SSEG_FUSE_BNFIN", "0") == "1"
        self._branch_streams = {}
        self._open_branches = []   # branches forked since the last join (build-time bookkeeping)

        self.convs, self.bns = {}, {}
        for m in self._modules():
            if isinstance(m, nn.Conv2d):
                if m.groups != 1:
                    continue   # depthwise convolutions (MobileNetV2) read their fp32 weight directly (csrc/depthwise.cu)
                self.convs[id(m)] = ConvW(m)
            elif isinstance(m, _BatchNorm):
                self.bns[id(m)] = BNS(m)
        self._alloc_params()
        self.img = torch.zeros(img_shape, device=self.dev, dtype=torch.float32) if part != "decoder" else None
        self._build_forward()
        if self.deferred_running:
            mods = list(self.deferred_running)

            def refresh_running():
                for m in mods:
                    ops.bn_running_from_tmp(m._tmp_running_mean, m._tmp_running_var, m._running_iter, m.running_mean,
                                            m.running_var)
            self.fwd.append(self.on_side(refresh_running))
            if not self.with_grad:
                self.fwd.append(self.join_side)
        if self.with_grad:
            self._build_backward()
        elif self.peer is not None:
            # forward-only synchronised steps have no gradient-bucket all-reduce to separate them: a one-float
            # all-reduce keeps a fast rank from re-zeroing statistics a slow peer is still reading
            fence = torch.zeros(1, device=self.dev)
            self.fwd.append(lambda: self.dist.all_reduce(fence))

    def _modules(self):
        mods = []
        for net in (self.enc, self.dec):
            if net is not None:
                mods += list(net.modules())
        return mods

    # ------------------------------------------------------------------------------------------ buffers
    def _alloc_params(self):
        dev = self.dev
        nf = sum(c.O * c.T * c.I for c in self.convs.values())
        nff = sum(c.O * c.ldf for c in self.convs.values())
        nd = sum(c.I * c.T * c.Opad for c in self.convs.values())
        self.wf_flat = torch.zeros(nff, device=dev, dtype=torch.bfloat16)
        self.wd_flat = torch.zeros(nd if self.with_grad else 0, device=dev, dtype=torch.bfloat16)
        small = sum(2 * b.Cp for b in self.bns.values()) + sum(_pad(c.O, 4) for c in self.convs.values()
                                                               if c.mod.bias is not None)
        self.g_small = small
        self.gflat = torch.zeros((small + nf) if self.with_grad else 0, device=dev, dtype=torch.float32)
        ns = sum(_pad(2 * b.C + 1, 4) + b.Cp for b in self.bns.values()) + 8  # + Cp: raw sum g'*y of the fused dgrad
        # with several ranks the statistics live in a peer-mapped arena (SyncBN without NCCL calls, csrc/peer.cu):
        # [per-BN sum|sqsum|count ... loss accumulators | per-BN s1|s2 partials ... | flags (int) | step (int)]
        self.peer = None
        npart = sum(2 * b.Cp for b in self.bns.values())
        import os
        if self.dist is not None and self.training and os.environ.get("SSEG_PEER_SYNC", "1") != "0":
            # ONE arena (and one step counter, at its end) per module, shared by the programs of every input shape: its
            # layout depends on the BatchNorm channel counts only. Ranks of a job draw different batch shapes (the
            # reference's loader), so they build / reuse programs at different times - a collective in program
            # construction, or handshake flags living in per-shape arenas, would hang or mis-pair them. The arena is
            # created with the FIRST program (step 1, all ranks together); later programs only look it up.
            from .peer import PeerArena
            nflag = 16 * len(self.bns) + 16
            # push-protocol inboxes behind the flags (csrc/peer.cuh: 8-byte {value, step} messages, one slot range per
            # sender): forward 2C+1 messages per layer and sender, backward 2Cp
            world = self.dist.get_world_size()
            ninbox = sum(2 * world * (_pad(2 * b.C + 1, 2) + 2 * b.Cp) for b in self.bns.values())
            need = _pad(ns + npart + nflag, 4) + ninbox
            owner = self.seg if self.seg is not None else (self.enc if self.enc is not None else self.dec)
            shared = owner.__dict__.get("_b200_peer")
            if shared is None:
                try:
                    self.peer = PeerArena(need, self.dist, dev)
                    ok = 1.0
                except Exception as exc:  # e.g. CUDA IPC unavailable between the ranks
                    self.peer, ok = None, 0.0
                    import warnings
                    warnings.warn("peer-memory SyncBN unavailable (%s); using NCCL all-reduces" % exc)
                agree = torch.tensor([ok], device=dev)
                self.dist.all_reduce(agree, op=self.dist.ReduceOp.MIN)  # every rank takes the same path
                if agree.item() < 1.0:
                    self.peer = None
                if self.peer is not None:
                    self.peer.floats[:ns + npart].zero_()
                owner.__dict__["_b200_peer"] = (self.peer, need)
            else:
                self.peer, have = shared
                assert have == need, "the module's BatchNorm layers changed after its first step program was built"
        if self.peer is not None:
            self.sflat = self.peer.floats[:ns + npart]
        else:
            self.sflat = torch.zeros(ns, device=dev, dtype=torch.float32)
        self.sinit = torch.zeros(self.sflat.numel(), device=dev, dtype=torch.float32)
        nv = sum(4 * b.Cp for b in self.bns.values())
        self.vflat = torch.zeros(nv, device=dev, dtype=torch.float32)
        of = od = 0
        og, ogs = small, 0
        for c in self.convs.values():
            c.wf = self.wf_flat[of:of + c.O * c.ldf].view(c.O, c.ldf)[:, :c.T * c.I]
            of += c.O * c.ldf
            if self.with_grad:
                c.wd = self.wd_flat[od:od + c.I * c.T * c.Opad].view(c.I, c.T * c.Opad)
                od += c.I * c.T * c.Opad
                c.gw = self.gflat[og:og + c.O * c.T * c.I].view(c.O, c.T * c.I)
                og += c.O * c.T * c.I
                if c.mod.bias is not None:
                    c.gb = self.gflat[ogs:ogs + c.O]
                    ogs += _pad(c.O, 4)
        os_, ov = 0, 0
        op_, ofl = ns, ns + npart
        oin = _pad(ns + npart + 16 * len(self.bns) + 16, 4)   # inboxes: behind the flags and the step counter
        for b in self.bns.values():
            b.stats = self.sflat[os_:os_ + 2 * b.C + 1]
            b.stats_off = os_
            # the padding slot behind [sum | sqsum | count] doubles as the grid-barrier counter of the fused conv+BN kernel
            # (an all-zero float is an all-zero uint32; the per-step copy from `sinit` re-zeroes it)
            b.counter = self.sflat[os_ + 2 * b.C + 1:os_ + 2 * b.C + 2] if (2 * b.C + 1) % 4 != 0 else None
            b.counter_bwd = self.sflat[os_ + 2 * b.C + 2:os_ + 2 * b.C + 3] if (2 * b.C + 1) % 4 == 1 else None
            os_ += _pad(2 * b.C + 1, 4)
            b.s2y = self.sflat[os_:os_ + b.Cp]
            os_ += b.Cp
            if self.peer is not None:
                b.part = self.sflat[op_:op_ + 2 * b.Cp]  # [s1 | s2] partial sums of the backward pass
                b.part_off, b.flag_off = op_, ofl
                op_ += 2 * b.Cp
                ofl += 16
                b.inbox_off = oin
                oin += 2 * self.world * _pad(2 * b.C + 1, 2)
                b.inbox_bwd_off = oin
                oin += 2 * self.world * 2 * b.Cp
                b.tot = torch.zeros(2 * b.Cp + 1, device=dev, dtype=torch.float32)  # s1_tot | s2_tot | pooled count
            b.mean, b.invstd, b.scale, b.shift = (self.vflat[ov + i * b.Cp: ov + (i + 1) * b.Cp] for i in range(4))
            ov += 4 * b.Cp
            if self.with_grad:
                b.dgamma = self.gflat[ogs:ogs + b.Cp]
                b.dbeta = self.gflat[ogs + b.Cp:ogs + 2 * b.Cp]
                ogs += 2 * b.Cp
        self.acc_main = self.sflat[os_:os_ + 4]
        self.acc_ds = self.sflat[os_ + 4:os_ + 8]
        if self.peer is not None:
            self.peer_step = self.peer.ints[ofl:ofl + 1]
        self.out = torch.zeros(2, device=dev, dtype=torch.float32)  # (loss, acc)

    def _new(self, *shape, dtype=torch.bfloat16, zero=False):
        f = torch.zeros if zero else torch.empty
        return f(shape, device=self.dev, dtype=dtype)

    def _new_act(self, n, h, w, c):
        """Activation with its channel count rounded up to 8 in storage (zero pad channels)."""
        cp = _pad(c, 8)
        buf = self._new(n, h, w, cp, zero=(cp != c))
        return Act(buf[..., :c] if cp != c else buf, buf)

    def _bn_mode(self, bns):
        if not self.training or not bns.mod.training:
            return ops.BN_EVAL
        sync = self.dist is not None or getattr(bns.mod, "_is_parallel", False)
        return ops.BN_TRAIN_SYNC if sync else ops.BN_TRAIN

    # ------------------------------------------------------------------------------------------ forward pieces
    def _wentry(self, c, wf=True, wd=True):
        return dict(w=c.mod.weight.detach(), wf=c.wf if wf else None, wd=c.wd if (wd and self.with_grad) else None,
                    g_src=None, g_dst=None, O=c.O, I=c.I, T=c.T, o_pad=c.Opad, channels_last=c.cl)

    def _prep_weights(self):
        """One launch re-lays out every conv weight (fp32 OIHW master -> bf16 GEMM operands)."""
        convs = [c for c in self.convs.values() if c.I != 3]  # the stem conv reads the fp32 master weight directly
        # parameter gradients are views of the flat gradient buffer: channels-last masters have the GEMM's [O][tap * I]
        # layout (see ConvW), the stem kernel accumulates OIHW directly
        for c in self.convs.values():
            if not self.with_grad:
                c.pg = None
            elif c.cl:
                k = c.k
                c.pg = c.gw.as_strided((c.O, c.I, k, k), (c.T * c.I, 1, k * c.I, c.I))
            else:
                c.pg = c.gw.view_as(c.mod.weight)
        self._late_convs, self._late_pending = set(), False
        if not self.overlap_relayout:
            if self.with_grad and self.split_prep:
                # (opt-in SSEG_SPLIT_PREP=1; measured slower than the single pass for every side-grid size, profiles/
                # r2_summary.md section 8) the data-gradient operands (the transposing half of the pass) are first read
                # ~2 ms later, by the backward pass: they are produced on the side stream while the forward pass runs, only
                # the forward operands (a plain cast of the channels-last masters) stay in front of the stem
                self.wtable = ops.WeightTable([self._wentry(c, wd=False) for c in convs], self.dev)
                self.wtable_d = ops.WeightTable([self._wentry(c, wf=False) for c in convs], self.dev)
                self.fwd.append(self.wtable.prep)
                # ... on a THIN grid: behind a launch with thousands of pending blocks the main stream's next kernels would
                # wait until its last block has been dispatched (measured: the stem started only when the pass had ended)
                nblk = int(os.environ.get("SSEG_PREP_SIDE_BLOCKS", "32"))
                self.fwd.append(self.on_side(lambda: self.wtable_d.prep(max_blocks=nblk)))
                return
            self.wtable = ops.WeightTable([self._wentry(c) for c in convs], self.dev)
            self.fwd.append(self.wtable.prep)
            return
        # early = the first ~2 M weight elements in execution (= module) order: stem, layer1, layer2 of a ResNet
        early, acc = [], 0
        for c in convs:
            acc += c.O * c.T * c.I
            if acc > 2_000_000 and early:
                break
            early.append(c)
        late = convs[len(early):]
        self._late_convs = {id(c) for c in late}
        t_early = ops.WeightTable([self._wentry(c, wd=False) for c in early], self.dev)
        t_late = ops.WeightTable([self._wentry(c, wd=False) for c in late], self.dev) if late else None
        t_dgrad = ops.WeightTable([self._wentry(c, wf=False) for c in convs], self.dev) if self.with_grad else None
        self._prep_tables = (t_early, t_late, t_dgrad)
        self.fwd.append(t_early.prep)

        def rest():
            if t_late is not None:
                t_late.prep()
            if not getattr(self, "serial", False):
                self._ev_late = torch.cuda.Event()
                self._ev_late.record(torch.cuda.current_stream(self.dev))
            if t_dgrad is not None:
                t_dgrad.prep()
        self.fwd.append(self.on_side(rest))
        self._late_pending = t_late is not None

    def _need_weights(self, cw):
        """Called before a convolution is scheduled: the first one whose forward operand is produced on the side stream
        makes the main stream wait for that (and only that) part of the side stream's work."""
        if self._late_pending and id(cw) in self._late_convs:
            self._late_pending = False

            def wait():
                if not getattr(self, "serial", False):
                    torch.cuda.current_stream(self.dev).wait_event(self._ev_late)
            self.fwd.append(wait)

    def _conv_geom(self, srcs, cw):
        """Input-side geometry of conv `cw` over the (virtual concat of) NHWC tensors `srcs` -> (geom, out H, out W)."""
        n, h, w, _ = srcs[0].shape
        if cw.stride == 1:
            geom = ops.make_geom(srcs, ops.conv_taps(cw.k, cw.dil))
            ho, wo = h, w
        else:
            assert cw.stride == 2 and cw.dil == 1 and len(srcs) == 1 and h % 2 == 0 and w % 2 == 0, \
                "stride-2 convolutions need even spatial sizes"
            dh, dw, src = ops.conv_s2_taps(cw.k)
            planes = ops.parity_planes(srcs[0])
            geom = ops.make_geom(planes, (dh, dw), tap_src=src, tap_koff=[t * cw.I for t in range(cw.T)])
            self.keep.append(planes)
            ho, wo = h // 2, w // 2
        self.keep.append(geom)
        return geom, ho, wo

    def conv_bn(self, xs, conv_mod, bn_mod, relu=True, res=None, chanmul=None, apply=True, post_add=None):
        """conv -> BN(train/eval) -> (+res) -> ReLU -> (*chanmul) (+post_add).  xs: Act or list of Acts (virtual concat).
        res: None | Act (identity shortcut) | ConvBNRec built with apply=False (projection shortcut).
        post_add: Act added AFTER the ReLU (FPN top-down path, models/models.py:561-563)."""
        xs = xs if isinstance(xs, list) else [xs]
        for a in xs + [r for r in (res, post_add) if isinstance(r, Act)]:
            a.uses += 1
        rec = ConvBNRec(self, xs, self.convs[id(conv_mod)], self.bns[id(bn_mod)], relu, res, chanmul, apply, post_add)
        self.records.append(rec)
        return rec if not apply else rec.a

    def _build_forward(self):
        from ..models import models as M
        from ..models import resnet as R
        N, H, W = self.N, self.H, self.W
        self._prep_weights()
        self.fwd.append(lambda: self.sflat.copy_(self.sinit))
        if self.peer is not None:
            self.fwd.append(lambda: ops.peer_step(self.peer_step))
        if self.part == "decoder":
            # module-level call decoder(conv_out): fp32 NCHW feature maps in -> engine layout
            self.feat_in, feats = [], []
            for shp in self.feat_shapes:
                n, c, h, w = shp
                src = self._new(n, c, h, w, dtype=torch.float32)
                act = Act(self._new(n, h, w, c))
                self.fwd.append(lambda src=src, act=act: ops.nchw_f32_to_nhwc_bf16(src, act.t))
                self.feat_in.append(src)
                feats.append(act)
        else:
            feats = self._build_encoder(R)
        self.feats = feats
        if self.part == "encoder":
            # module-level call encoder(x, return_feature_maps=True): hand the maps back as fp32 NCHW tensors
            self.feat_out = []
            for f in feats:
                parts = f if isinstance(f, list) else [f]   # HRNet hands back one map = concat of its 4 branches
                n, h, w, _ = parts[0].t.shape
                dst = self._new(n, sum(q.t.shape[3] for q in parts), h, w, dtype=torch.float32)
                off = 0
                for q in parts:
                    c = q.t.shape[3]
                    if len(parts) == 1:
                        self.fwd.append(lambda q=q, dst=dst: ops.nhwc_bf16_to_nchw_f32(q.t, dst))
                    else:
                        tmp = self._new(n, c, h, w, dtype=torch.float32)
                        self.fwd.append(lambda q=q, tmp=tmp, sl=dst[:, off:off + c]: (ops.nhwc_bf16_to_nchw_f32(q.t, tmp),
                                                                                    sl.copy_(tmp)))
                    off += c
                self.feat_out.append(dst)
            return
        # ---- decoder
        dec = self.dec
        self.logits_ds = None
        p_drop_main = p_drop_ds = 0.0
        if not isinstance(dec, (M.C1, M.C1DeepSup)) and any(isinstance(f, list) for f in feats):
            raise NotImplementedError("HRNetV2 features (a virtual concat of 4 branches) feed the C1 decoders only")
        if isinstance(dec, (M.PPM, M.PPMDeepsup)):
            conv5 = feats[-1]
            n, h, w, c5 = conv5.t.shape
            srcs = [conv5]
            for bi, (scale, branch) in enumerate(zip(dec.pool_scales, dec.ppm)):
                with self.branch(bi):
                    pr = AvgPoolRec(self, conv5, scale)
                    self.records.append(pr)
                    y = self.conv_bn(pr.a, branch[1], branch[2])
                    ur = UpsampleRec(self, y, h, w)
                    self.records.append(ur)
                srcs.append(ur.a)
            self.join_branches()
            drop = dec.conv_last[3]
            p_drop_main = drop.p if (self.training and drop.training) else 0.0
            self.mask_main = self._new(n, 512, dtype=torch.float32) if p_drop_main > 0 else None
            x = self.conv_bn(srcs, dec.conv_last[0], dec.conv_last[1], chanmul=self.mask_main)
            cls = ClassifierRec(self, x, self.convs[id(dec.conv_last[4])])
            self.records.append(cls)
            self.logits = cls.logits
            if isinstance(dec, M.PPMDeepsup) and not self.inference:
                p_drop_ds = dec.dropout_deepsup.p if (self.training and dec.dropout_deepsup.training) else 0.0
                self.mask_ds = self._new(n, dec.cbr_deepsup[0].out_channels, dtype=torch.float32) if p_drop_ds > 0 else None
                y = self.conv_bn(feats[-2], dec.cbr_deepsup[0], dec.cbr_deepsup[1], chanmul=self.mask_ds)
                cls2 = ClassifierRec(self, y, self.convs[id(dec.conv_last_deepsup)])
                self.records.append(cls2)
                self.logits_ds = cls2.logits
        elif isinstance(dec, (M.C1, M.C1DeepSup)):
            x = self.conv_bn(feats[-1], dec.cbr[0], dec.cbr[1])
            cls = ClassifierRec(self, x, self.convs[id(dec.conv_last)])
            self.records.append(cls)
            self.logits = cls.logits
            if isinstance(dec, M.C1DeepSup) and not self.inference:
                y = self.conv_bn(feats[-2], dec.cbr_deepsup[0], dec.cbr_deepsup[1])
                cls2 = ClassifierRec(self, y, self.convs[id(dec.conv_last_deepsup)])
                self.records.append(cls2)
                self.logits_ds = cls2.logits
        elif isinstance(dec, M.UPerNet):
            # reference models/models.py:543-586
            conv5 = feats[-1]
            n, h, w, c5 = conv5.t.shape
            srcs = [conv5]
            for scale, branch in zip(dec.pool_scales, dec.ppm_conv):
                pr = AvgPoolRec(self, conv5, scale)
                self.records.append(pr)
                ur = UpsampleRec(self, pr.a, h, w)       # here the 1x1 conv comes AFTER the up-sampling (:548-552)
                self.records.append(ur)
                srcs.append(self.conv_bn(ur.a, branch[0], branch[1]))
            f = self.conv_bn(srcs, dec.ppm_last_conv[0], dec.ppm_last_conv[1])
            levels = [f]
            for i in reversed(range(len(feats) - 1)):
                lat_in = feats[i]
                _, hi, wi, _ = lat_in.t.shape
                up = UpsampleRec(self, f, hi, wi)          # top-down branch (:560-561)
                self.records.append(up)
                f = self.conv_bn(lat_in, dec.fpn_in[i][0], dec.fpn_in[i][1], post_add=up.a)   # lateral + add (:557-563)
                levels.append(self.conv_bn(f, dec.fpn_out[i][0][0], dec.fpn_out[i][0][1]))
            levels.reverse()                                # [P2 .. P5]
            _, h2, w2, _ = levels[0].t.shape
            fusion = [levels[0]]
            for lv in levels[1:]:
                ur = UpsampleRec(self, lv, h2, w2)
                self.records.append(ur)
                fusion.append(ur.a)
            x = self.conv_bn(fusion, dec.conv_last[0][0], dec.conv_last[0][1])
            cls = ClassifierRec(self, x, self.convs[id(dec.conv_last[1])])
            self.records.append(cls)
            self.logits = cls.logits
        else:
            raise NotImplementedError("decoder %s is not built on the B200 engine yet" % type(dec).__name__)
        self.p_drop = (p_drop_main, p_drop_ds)
        self.num_class = cls.cw.O
        # dropout masks are drawn at the START of the step (they only depend on the RNG)
        if p_drop_main > 0 or p_drop_ds > 0:
            self.fwd.insert(0, self._draw_masks)
        # ---- head
        C8 = _pad(self.num_class, 8)
        if self.part == "decoder":
            # module-level call: probabilities at seg_size (use_softmax) or log-probabilities at feature resolution
            n, h, w, _ = self.logits.shape
            if self.inference:
                hs, ws = self.seg_size
                self.probs = self._new(n, self.num_class, hs, ws, dtype=torch.float32)
                self.fwd.append(lambda: ops.upsample_softmax(self.logits[..., :C8], self.num_class, self.probs))
                self.outputs = [self.probs]
            else:
                self.outputs = []
                for lg in [self.logits] + ([self.logits_ds] if self.logits_ds is not None else []):
                    o = self._new(n, self.num_class, h, w, dtype=torch.float32)
                    self.fwd.append(lambda lg=lg, o=o: ops.upsample_softmax(lg[..., :C8], self.num_class, o,
                                                                             log_output=True))
                    self.outputs.append(o)
        elif self.inference:
            hs, ws = self.seg_size
            if self.head_out is not None:
                # multi-scale evaluation (eval.py:63-72): this scale ADDS weight * softmax into a shared score map
                assert tuple(self.head_out.shape) == (self.N, self.num_class, hs, ws)
                self.probs = self.head_out
            else:
                self.probs = self._new(self.N, self.num_class, hs, ws, dtype=torch.float32)
            wgt, acc = self.head_weight, self.head_out is not None
            self.fwd.append(lambda: ops.upsample_softmax(self.logits[..., :C8], self.num_class, self.probs, weight=wgt,
                                                         accumulate=acc))
        else:
            n, h, w, _ = self.logits.shape
            self.label = torch.full((n, h, w), -1, device=self.dev, dtype=torch.int64)
            loss = LossRec(self)
            self.records.append(loss)

    def _residual_chain(self, x, blocks):
        """nn.Sequential of BasicBlock / Bottleneck: (conv, bn) stages, identity or projected shortcut, ReLU."""
        for block in blocks:
            inp = x
            res = inp
            if block.downsample is not None:
                res = self.conv_bn(inp, block.downsample[0], block.downsample[1], relu=False, apply=False)
            stages = block.stages()
            for i, (cv, bn) in enumerate(stages):
                last = i == len(stages) - 1
                x = self.conv_bn(x, cv, bn, relu=True, res=res if last else None)
        return x

    def _build_hrnet(self):
        """HRNetV2.forward / HighResolutionModule.forward (reference models/hrnet.py:395-437, :225-250)."""
        enc = self.enc
        assert self.H % 32 == 0 and self.W % 32 == 0, "HRNetV2 needs H, W multiples of 32 (five stride-2 levels)"
        stem = StemRec(self, self.convs[id(enc.conv1)], self.bns[id(enc.bn1)])
        self.records.append(stem)
        x = self.conv_bn(stem.a, enc.conv2, enc.bn2)
        x = self._residual_chain(x, enc.layer1)
        ys = [x]
        for si in (2, 3, 4):
            trans, stage = getattr(enc, "transition%d" % (si - 1)), getattr(enc, "stage%d" % si)
            xs = []
            for i, tr in enumerate(trans):
                if i < len(ys):
                    if tr is None:
                        xs.append(ys[i])
                    else:  # the reference feeds y_list[-1] here from stage 3 on (hrnet.py:414-430); None for W48
                        xs.append(self.conv_bn(ys[-1] if si > 2 else ys[i], tr[0], tr[1]))
                else:
                    t = ys[-1]
                    for link in tr:
                        t = self.conv_bn(t, link[0], link[1])
                    xs.append(t)
            for module in stage:
                xs = self._hr_module(module, xs)
            ys = xs
        _, h, w, _ = ys[0].t.shape
        srcs = [ys[0]]
        for y in ys[1:]:
            ur = UpsampleRec(self, y, h, w)
            self.records.append(ur)
            srcs.append(ur.a)
        return [srcs]   # one feature map: the virtual concat of the four branches (torch.cat at hrnet.py:434)

    def _hr_module(self, module, xs):
        nb = module.num_branches
        outs_ = []
        for i in range(nb):
            with self.branch(i):
                outs_.append(self._residual_chain(xs[i], module.branches[i]))
        self.join_branches()
        xs = outs_
        if module.fuse_layers is None:
            return xs
        outs = []
        for i, row in enumerate(module.fuse_layers):
            # the rows of an exchange unit only READ the branch outputs: independent in the forward pass. Their backward
            # passes accumulate into the same branch gradients, so they stay in order on the main stream.
            with self.branch(i, backward=False):
                terms = []
                for j in range(nb):
                    if j == i:
                        terms.append(("id", xs[j]))
                    elif j > i:   # 1x1 conv + BN at the low resolution; the up-sampling happens inside the sum kernel
                        terms.append(("up", self.conv_bn(xs[j], row[j][0], row[j][1], relu=False, apply=False)))
                    else:         # chain of stride-2 3x3 convs, ReLU between the links only
                        t, links = xs[j], list(row[j])
                        for link in links[:-1]:
                            t = self.conv_bn(t, link[0], link[1])
                        terms.append(("same", self.conv_bn(t, links[-1][0], links[-1][1], relu=False, apply=False)))
                sr = SumRec(self, terms)
                self.records.append(sr)
            outs.append(sr.a)
        self.join_branches()
        return outs

    def _build_mobilenet(self):
        """MobileNetV2Dilated.forward (reference models/models.py:312-323) over InvertedResidual blocks
        (models/mobilenet.py:38-76), inference only: every BatchNorm (running statistics) and ReLU6 is folded into the
        kernel that produces the tensor — 1x1 convolutions on the implicit-GEMM kernel (sseg_conv_igemm_affine, shortcut as
        addend), depthwise 3x3 convolutions and the 3-channel first layer on their own kernels (csrc/depthwise.cu)."""
        enc = self.enc
        if self.with_grad or self.training:
            raise NotImplementedError("MobileNetV2 runs on the B200 engine in eval mode only (no ReLU6 / depthwise backward)")

        def affine(bn_mod):
            bns = self.bns[id(bn_mod)]
            if self._bn_mode(bns) != ops.BN_EVAL:
                raise NotImplementedError("MobileNetV2 runs on the B200 engine in eval mode only")
            _emit_bn_forward(self, bns, ops.BN_EVAL, 1, None, None, False, None, None, None, None)
            return bns.scale, bns.shift
        first = enc.features[0]
        conv0, bn0 = first[0], first[1]
        assert conv0.in_channels == 3 and conv0.stride == (2, 2) and conv0.kernel_size == (3, 3)
        sc, sh = affine(bn0)
        ho, wo = (self.H - 1) // 2 + 1, (self.W - 1) // 2 + 1
        x = Act(self._new(self.N, ho, wo, conv0.out_channels))
        w0 = conv0.weight.detach()
        self.fwd.append(lambda x=x, sc=sc, sh=sh: ops.stem_conv_affine(self.img, w0, x.t, scale=sc, shift=sh, relu6=True))
        feats = []
        self.block_outs = [x]   # per-block outputs (tools / tests compare them with the oracle's)
        for idx in range(1, enc.total_idx):
            block = enc.features[idx]
            inp = x
            stages = block.stages()
            for si, (cv, bn, act) in enumerate(stages):
                last = si == len(stages) - 1
                sc, sh = affine(bn)
                n, h, w, c = x.t.shape
                if cv.groups != 1:
                    assert cv.groups == c == cv.out_channels and cv.kernel_size == (3, 3) and cv.padding[0] == cv.dilation[0]
                    s_, d_ = cv.stride[0], cv.dilation[0]
                    y = Act(self._new(n, (h - 1) // s_ + 1, (w - 1) // s_ + 1, c))
                    wdw = cv.weight.detach()
                    self.fwd.append(lambda x=x, y=y, wdw=wdw, s_=s_, d_=d_, sc=sc, sh=sh, act=act:
                                    ops.dwconv_affine(x.t, wdw, y.t, stride=s_, dilation=d_, scale=sc, shift=sh, relu6=act))
                else:
                    cw = self.convs[id(cv)]
                    assert cw.k == 1 and cw.stride == 1
                    geom, _, _ = self._conv_geom([x.t], cw)
                    y = self._new_act(n, h, w, cw.O)
                    relu = (ops.RELU_AFTER_ADD | ops.RELU6) if act else ops.RELU_NONE
                    add = inp.tp if (last and block.use_res_connect) else None   # x + conv(x), no activation after the add
                    self._need_weights(cw)
                    self.fwd.append(lambda geom=geom, cw=cw, y=y, sc=sc, sh=sh, relu=relu, add=add:
                                    ops.conv_igemm_affine(geom, cw.wf, cw.O, y.tp, sc, sh, relu=relu, addend=add))
                x = y
            self.block_outs.append(x)
            if idx in enc.down_idx:
                feats.append(x)
        feats.append(x)
        return feats

    def _build_encoder(self, R):
        enc = self.enc
        from ..models import hrnet as HR
        from ..models import models as M_
        if isinstance(enc, HR.HRNetV2):
            return self._build_hrnet()
        if isinstance(enc, M_.MobileNetV2Dilated):
            return self._build_mobilenet()
        # ---- stem (reference models/resnet.py:100-109, models/models.py:256-259)
        stem = StemRec(self, self.convs[id(enc.conv1)], self.bns[id(enc.bn1)])
        self.records.append(stem)
        x = stem.a
        x = self.conv_bn(x, enc.conv2, enc.bn2)
        x = self.conv_bn(x, enc.conv3, enc.bn3)
        mp = MaxPoolRec(self, x)
        self.records.append(mp)
        x = mp.a
        feats = []
        for layer in (enc.layer1, enc.layer2, enc.layer3, enc.layer4):
            x = self._residual_chain(x, layer)
            feats.append(x)
        return feats

    def _draw_masks(self):
        for name, mask, p in (("main", getattr(self, "mask_main", None), self.p_drop[0]),
                              ("deepsup", getattr(self, "mask_ds", None), self.p_drop[1])):
            if mask is None:
                continue
            if self.injected_masks is not None and name in self.injected_masks:
                mask.copy_(self.injected_masks[name].to(mask.dtype) / (1.0 - p))
            else:
                # nn.Dropout2d (models/models.py:460,464): one Bernoulli(1-p) draw per (n, c), survivors scaled by 1/(1-p)
                mask.uniform_().ge_(p).mul_(1.0 / (1.0 - p))

    # ------------------------------------------------------------------------------------------ backward
    def _build_backward(self):
        self.bwd.append(lambda: self.gflat.zero_())
        self._bucket_works = []
        # The data-parallel gradient bucket (reference: backward of nn.DataParallel's Broadcast) as FOUR
        # NCCL all-reduces of contiguous slices of the flat fp32 gradient buffer, issued in backward order on the side
        # stream right behind the weight-gradient GEMMs that fill them, so they overlap the rest of the backward pass:
        #   [ small (BN, biases) | encoder stem..layer2 | encoder layer3 | encoder layer4 | decoder ]
        buckets = []
        if self.overlap_relayout or self.split_prep:
            self.bwd.append(self.join_side)   # the data-gradient operands were produced on the side stream
        if (self.dist is not None or self.overlap_relayout) and self.enc is not None and self.dec is not None:
            dec_ids = {id(m) for m in self.dec.modules()}
            late = self.enc.layer4 if hasattr(self.enc, "layer4") else self.enc.stage4   # the encoder's last stage
            l4_ids = {id(m) for m in late.modules()}
            off = lambda ids: min(c.gw.storage_offset() for c in self.convs.values() if id(c.mod) in ids)
            dec_off, l4_off = off(dec_ids), off(l4_ids)
            assert l4_off < dec_off
            end = self.gflat.numel()
            buckets = [(dec_ids, self.gflat[dec_off:end]), (dec_ids | l4_ids, self.gflat[l4_off:dec_off])]
            rest = self.gflat[:l4_off]
            if hasattr(self.enc, "layer4") and hasattr(self.enc, "layer3") and self.dist is not None:
                # a fourth slice (layer3: 28 MB of ResNet50's remaining 34 MB) keeps the all-reduce that is exposed at the
                # very end of the step small
                l3_ids = {id(m) for m in self.enc.layer3.modules()}
                l3_off = off(l3_ids)
                if 0 < l3_off < l4_off:
                    buckets.append((dec_ids | l4_ids | l3_ids, self.gflat[l3_off:l4_off]))
                    rest = self.gflat[:l3_off]
        # 1 / world_size (= the reference's mean over per-GPU losses, train.py:42) is applied ONCE, to the loss gradient
        # (LossRec.backward): every gradient of the step then comes out already divided, the bucket all-reduces sum them
        gtables = None
        pending = list(buckets)
        nclosed = 0
        for rec in reversed(self.records):
            mod = getattr(getattr(rec, "cw", None), "mod", None)
            if pending and mod is not None and id(mod) not in pending[0][0]:
                # every record of this bucket has emitted its weight gradient: reduce it behind them on the side stream
                sl = pending.pop(0)[1]
                gt = gtables[nclosed] if gtables is not None else None
                nclosed += 1

                def close_bucket(sl=sl, gt=gt):
                    if self.dist is not None:
                        # asynchronous: the collective runs on the communicator's own stream behind the side stream's
                        # work so far, the weight-gradient GEMMs enqueued after it do NOT wait for it (a synchronous
                        # call parks the side stream until the all-reduce has finished); joined at the end of the step
                        self._bucket_works.append(self.dist.all_reduce(sl, async_op=True))
                self.bwd.append(self.on_side(close_bucket))
            k = getattr(rec, "branch", None) if self.use_branches else None
            if isinstance(rec, AvgPoolRec):
                k = None   # the pyramid's pools share ONE grouped backward kernel: it runs on the main stream
            b0 = len(self.bwd)
            rec.backward()
            if len(self.bwd) == b0:
                continue   # nothing emitted (e.g. a record driven by its consumer): no ordering point
            if k is None:
                if self._open_branches:   # main-stream work that may read what the branches produced: join first
                    joins = [self._join(j) for j in self._open_branches]
                    self._open_branches = []
                    self.bwd[b0:b0] = joins
            else:
                seg = self.bwd[b0:]
                self.bwd[b0:] = ([] if k in self._open_branches else [self._fork(k)]) + [self._on_branch(k, f) for f in seg]
                if k not in self._open_branches:
                    self._open_branches.append(k)
        self.join_branches(into=self.bwd)
        self.bwd.append(self.join_side)
        if self.dist is not None:
            def join_buckets():
                for w in self._bucket_works:
                    w.wait()       # NCCL: the current stream waits for the communicator's stream (capturable)
                del self._bucket_works[:]
            self.bwd.append(join_buckets)
            if buckets:
                for _, sl in pending:  # (degenerate nets: buckets never closed)
                    self.bwd.append(lambda sl=sl: self.dist.all_reduce(sl))
                self.bwd.append(lambda: self.dist.all_reduce(rest))
            else:
                self.bwd.append(lambda: self.dist.all_reduce(self.gflat))
        self._pg = {id(c): c.pg for c in self.convs.values()}

    def on_side(self, fn):
        """Closure that runs `fn` on the side stream, ordered after everything enqueued so far on the main stream."""
        def run():
            if getattr(self, "serial", False):
                return fn()
            main = torch.cuda.current_stream(self.dev)
            ev = torch.cuda.Event()
            ev.record(main)
            self.side.wait_event(ev)
            with torch.cuda.stream(self.side):
                fn()
        return run

    def join_side(self):
        if not getattr(self, "serial", False):
            torch.cuda.current_stream(self.dev).wait_stream(self.side)

    # ---- branch streams (fork / run / join closures; all no-ops with prog.serial, like the side stream)
    def _bstream(self, k):
        s = self._branch_streams.get(k)
        if s is None:
            s = self._branch_streams[k] = torch.cuda.Stream(self.dev)
        return s

    def _fork(self, k):
        def run():
            if getattr(self, "serial", False):
                return
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.dev))
            self._bstream(k).wait_event(ev)
        return run

    def _on_branch(self, k, fn):
        def run():
            if getattr(self, "serial", False):
                return fn()
            with torch.cuda.stream(self._bstream(k)):
                fn()
        return run

    def _join(self, k):
        def run():
            if not getattr(self, "serial", False):
                torch.cuda.current_stream(self.dev).wait_stream(self._bstream(k))
        return run

    def branch(self, k, backward=True):
        """Context manager for the FORWARD schedule: closures and records created inside run on branch stream k
        (forked from the main stream at the point of entry). Call join_branches() before anything consumes them.
        backward=False: only the forward closures branch; the records' backward closures stay on the main stream (needed
        when different branches ACCUMULATE into the same gradient buffers)."""
        P = self

        class _Scope:
            def __enter__(self_):
                self_.f0, self_.r0 = len(P.fwd), len(P.records)

            def __exit__(self_, *exc):
                if not P.use_branches or exc[0] is not None:
                    return False
                seg = P.fwd[self_.f0:]
                if seg:
                    P.fwd[self_.f0:] = [P._fork(k)] + [P._on_branch(k, f) for f in seg]
                    if k not in P._open_branches:
                        P._open_branches.append(k)
                if backward:
                    for r in P.records[self_.r0:]:
                        r.branch = k
                return False
        return _Scope()

    def join_branches(self, into=None):
        """Append the joins of every open branch to `into` (default: the forward schedule)."""
        lst = self.fwd if into is None else into
        for k in self._open_branches:
            lst.append(self._join(k))
        self._open_branches = []

    def grad_target(self, act, shape_like=None):
        """(buffer, accumulate?) for writing a gradient contribution of `act`."""
        if act.g is None:
            act.g = torch.empty_like(act.tp if shape_like is None else shape_like)
        acc = act.gw
        act.gw = True
        act.gcount += 1
        return act.g, acc

    # ------------------------------------------------------------------------------------------ execution
    def load_features(self, feats):
        for dst, src in zip(self.feat_in, feats):
            dst.copy_(src, non_blocking=True)

    def load_inputs(self, img, label=None):
        self.img.copy_(img, non_blocking=True)
        if label is not None:
            self.label.copy_(label, non_blocking=True)

    def run_eager(self):
        if self.dry_run:
            raise RuntimeError("a dry-run program only describes the schedule; there is no CPU execution path")
        for f in self.fwd:
            f()
        for f in self.bwd:
            f()

    def capture(self, warm=True):
        """Capture the whole step into one CUDA graph. warm=True: two warm-up executions on a side stream first (lazy
        one-time initialisation - kernel attributes, tensor-map cache - must not happen inside the capture).
        warm=False: the caller has already run this program eagerly; nothing is executed here, so no SyncBN handshake /
        NCCL call is issued - under torch.distributed ranks may then capture at different steps (they see different batch
        shapes)."""
        if warm:
            # the warm-up executes the step: keep it side-effect free on the module (BN running statistics)
            bufs = [b for m in self._modules() if isinstance(m, _BatchNorm) for b in m.buffers(recurse=False)]
            saved = [b.clone() for b in bufs]
            s = torch.cuda.Stream(self.dev)
            s.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(s):
                self.run_eager()
                self.run_eager()
            torch.cuda.current_stream(self.dev).wait_stream(s)
            torch.cuda.synchronize(self.dev)
            for b, v in zip(bufs, saved):
                b.copy_(v)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.run_eager()
        self.graph = g
        return g

    def run(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self.run_eager()

    def num_launches(self):
        """Kernel launches of libsseg_b200 per step (counted by the library on one eager run)."""
        from . import _C
        _C.lib().sseg_launch_count_reset()
        self.run_eager()
        return int(_C.lib().sseg_launch_count())

    # ------------------------------------------------------------------------------------------ gradients
    def param_grads(self):
        """{parameter: fp32 gradient tensor in the parameter's own layout}. The tensors are the program's static
        buffers, rewritten by every run (conv weights: `grad_to_oihw` outputs; BN / bias: views of the flat buffer)."""
        out = {}
        for c in self.convs.values():
            out[c.mod.weight] = self._pg[id(c)]
            if c.mod.bias is not None:
                out[c.mod.bias] = c.gb
        for b in self.bns.values():
            if b.mod.weight is not None:
                out[b.mod.weight] = b.dgamma[:b.C]
                out[b.mod.bias] = b.dbeta[:b.C]
        return out


# ======================================================================================================= records
class StemRec:
    """conv1 (3->64, 3x3, stride 2) + bn1 + relu1: models/resnet.py:100-102, :153."""

    def __init__(self, P, cw, bns):
        self.P, self.cw, self.bns = P, cw, bns
        assert cw.I == 3 and cw.O == 64 and cw.k == 3 and cw.stride == 2, "deep-stem conv1 expected"
        N, H, W = P.N, P.H, P.W
        ho, wo = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
        self.y = P._new(N, ho, wo, 64)
        self.a = Act(P._new(N, ho, wo, 64))
        self.a.producer = self
        self.relu, self.res, self.post_add, self.chanmul, self.apply, self.fused = True, None, None, None, True, False
        self.dy_pre = None   # set when the consumer's data-gradient kernel already did this layer's BN backward
        self.mode = P._bn_mode(bns)
        self.count = N * ho * wo
        w = cw.mod.weight
        train = self.mode != ops.BN_EVAL
        st = bns.stats
        P.fwd.append(lambda: ops.stem_conv_fwd(P.img, w.detach(), self.y, st[:64] if train else None,
                                               st[64:128] if train else None))
        _emit_bn_forward(P, bns, self.mode, self.count, self.y, self.a.t, True, None, None, None, None)

    def backward(self):
        P, bns = self.P, self.bns
        if self.a.g is None:
            return
        if self.dy_pre is not None:
            dy = self.dy_pre
        else:
            dy = torch.empty_like(self.y)
            _emit_bn_backward(P, bns, self.mode, self.count, self.a.g, None, self.y, dy, None, None, mask_from_y=True,
                              fused=self.fused)
        gw = self.cw.gw
        P.bwd.append(P.on_side(lambda: ops.stem_conv_wgrad(P.img, dy, gw.view(64, 3, 3, 3))))


def _emit_bn_forward(P, bns, mode, count, y, out, relu, res, rscale, rshift, chanmul, res_after_relu=False, finalized=False):
    """finalized: the producing conv launch already wrote mean / inv_std / scale / shift (sseg_conv_igemm_bnfin)."""
    if finalized:
        if out is not None:
            P.fwd.append(lambda: ops.bn_apply(y, bns.scale, bns.shift, out, relu=relu, res=res, rscale=rscale, rshift=rshift,
                                              chanmul=chanmul, res_after_relu=res_after_relu))
        return
    m = bns.mod
    C, Cp = bns.C, bns.Cp
    st = bns.stats
    # finalize works on the C real channels; apply / backward kernels on the Cp-long (zero padded) vectors
    mean, invstd, scale, shift = bns.mean[:C], bns.invstd[:C], bns.scale[:C], bns.shift[:C]
    running = (m.running_mean, m.running_var, getattr(m, "_tmp_running_mean", None), getattr(m, "_tmp_running_var", None),
               getattr(m, "_running_iter", None))
    upd = mode != ops.BN_EVAL and m.track_running_stats and m.running_mean is not None
    mom = m.momentum if m.momentum is not None else 0.1
    w = m.weight.detach() if m.weight is not None else None
    b = m.bias.detach() if m.bias is not None else None
    if mode == ops.BN_TRAIN_SYNC:
        # local pixel count rides along with the sums so ranks with different batch shapes pool correctly
        P.sinit[bns.stats_off + 2 * C] = float(count)
    if mode == ops.BN_TRAIN_SYNC and P.peer is not None:
        cnt_out = bns.tot[2 * Cp:2 * Cp + 1]
        # running_mean / running_var (= accumulators / running_iter, batchnorm.py:136-137) are outputs of the step only: their
        # refresh is collected and runs on the side stream after the forward pass instead of as a second launch per layer
        # on the forward chain
        defer = upd and P.defer_running
        flag_or_inbox = bns.inbox_off if P.peer_ll else bns.flag_off
        P.fwd.append(lambda: ops.bn_finalize_peer(P.peer, bns.stats_off, flag_or_inbox, P.peer_step, w, b, m.eps, mom,
                                                  mean, invstd, scale, shift, cnt_out, running=running,
                                                  update_running=upd, defer_running=defer, ll=P.peer_ll))
        if defer:
            P.deferred_running.append(m)
    elif mode == ops.BN_TRAIN and out is not None and P.fuse_finalize and C == Cp:
        # single-GPU training: finalize fused into the apply kernel (one kernel boundary less per layer)
        rmean, rvar = (m.running_mean, m.running_var) if upd else (None, None)
        P.fwd.append(lambda: ops.bn_finalize_apply(st[:C], st[C:2 * C], count, w, b, m.eps, mom, bns.mean, bns.invstd,
                                                   bns.scale, bns.shift, y, out, relu=relu, res=res, rscale=rscale,
                                                   rshift=rshift, chanmul=chanmul, res_after_relu=res_after_relu,
                                                   running_mean=rmean, running_var=rvar))
        return
    else:
        if mode == ops.BN_TRAIN_SYNC and P.dist is not None:
            P.fwd.append(lambda: P.dist.all_reduce(st))
        cdev = st[2 * C:2 * C + 1] if mode == ops.BN_TRAIN_SYNC else None
        P.fwd.append(lambda: ops.bn_finalize(st[:C], st[C:2 * C], count, w, b, m.eps, mom, mode, mean, invstd,
                                             scale, shift, running=running, update_running=upd, count_dev=cdev))
    if out is not None:
        P.fwd.append(lambda: ops.bn_apply(y, bns.scale, bns.shift, out, relu=relu, res=res, rscale=rscale, rshift=rshift,
                                          chanmul=chanmul, res_after_relu=res_after_relu))


def _emit_bn_backward(P, bns, mode, count, g, a, y, dy, dres, chanmul, mask_from_y=False, fused=False):
    """g: gradient w.r.t. the layer output.  ReLU mask: `a` (saved output) if given, else recomputed from y when
    mask_from_y (layers without a shortcut), else the layer has no ReLU."""
    C = bns.Cp  # vector length of the backward kernels (= channels of g / y / dy in storage)
    st = bns.stats
    sc = bns.scale
    fs = bns.shift if (mask_from_y and a is None) else None
    if fused:
        # the consumer's dgrad epilogue already accumulated s1 (= dbeta) and the raw sum g'*y (sseg_conv_igemm_bnbwd, or,
        # for a layer with a shortcut - mask from its saved output `a`, shortcut gradient `dres` - sseg_conv_igemm_bnbwd_res)
        assert (a is None) != (fs is None) and chanmul is None and (dres is None or a is not None)
        if mode == ops.BN_TRAIN_SYNC and P.peer is not None and P.peer_fuse_bwd:
            # handshake + pooling over NVLink inside the apply pass (one dependent launch less per layer)
            cnt = bns.tot[2 * C:2 * C + 1]
            P.bwd.append(lambda: ops.bn_bwd_apply_peer(P.peer, bns.part_off, bns.flag_off + 8, P.peer_step, g, a, y, bns.mean,
                                                       bns.invstd, sc, cnt, dy, bns.dbeta, bns.dgamma, dres=dres, fshift=fs,
                                                       s2_raw=True))
        elif mode == ops.BN_TRAIN_SYNC and P.peer is not None:
            t1, t2, cnt = bns.tot[:C], bns.tot[C:2 * C], bns.tot[2 * C:2 * C + 1]
            fo = bns.inbox_bwd_off if P.peer_ll else bns.flag_off + 8
            P.bwd.append(lambda: ops.bn_bwd_peer_sum(P.peer, bns.part_off, fo, P.peer_step, t1, t2, bns.dbeta,
                                                     bns.dgamma, mean=bns.mean, invstd=bns.invstd, s2_raw=True,
                                                     ll=P.peer_ll))
            P.bwd.append(lambda: ops.bn_bwd_apply(g, a, y, bns.mean, bns.invstd, sc, t1, t2, count, dy, dres=dres,
                                                  count_dev=cnt, fshift=fs))
        else:
            P.bwd.append(lambda: ops.bn_bwd_apply(g, a, y, bns.mean, bns.invstd, sc, bns.dbeta, bns.s2y, count, dy, dres=dres,
                                                  eval_mode=(mode == ops.BN_EVAL), fshift=fs, s2_raw=True,
                                                  dgamma_out=bns.dgamma))
        return
    if mode == ops.BN_EVAL:
        P.bwd.append(lambda: ops.bn_bwd_apply(g, a, y, None, None, sc, None, None, 1.0, dy, dres=dres, chanmul=chanmul,
                                              eval_mode=True, fshift=fs))
        # dgamma / dbeta of a frozen BN still exist in the reference (affine params stay trainable under fix_bn)
        P.bwd.append(lambda: ops.bn_bwd_reduce(g, a, y, bns.mean, bns.invstd, bns.dbeta, bns.dgamma, chanmul=chanmul,
                                               scale=sc, fshift=fs))
        return
    if mode == ops.BN_TRAIN_SYNC and P.peer is not None:
        # partial sums into the peer-mapped arena -> handshake + pooled totals (csrc/peer.cu) -> apply
        p1, p2 = bns.part[:C], bns.part[C:2 * C]
        t1, t2, cnt = bns.tot[:C], bns.tot[C:2 * C], bns.tot[2 * C:2 * C + 1]
        P.bwd.append(lambda: ops.bn_bwd_reduce(g, a, y, bns.mean, bns.invstd, p1, p2, chanmul=chanmul, scale=sc, fshift=fs))
        if P.peer_fuse_bwd:
            P.bwd.append(lambda: ops.bn_bwd_apply_peer(P.peer, bns.part_off, bns.flag_off + 8, P.peer_step, g, a, y, bns.mean,
                                                       bns.invstd, sc, cnt, dy, bns.dbeta, bns.dgamma, dres=dres,
                                                       chanmul=chanmul, fshift=fs))
            return
        fo = bns.inbox_bwd_off if P.peer_ll else bns.flag_off + 8
        P.bwd.append(lambda: ops.bn_bwd_peer_sum(P.peer, bns.part_off, fo, P.peer_step, t1, t2, bns.dbeta,
                                                 bns.dgamma, ll=P.peer_ll))
        P.bwd.append(lambda: ops.bn_bwd_apply(g, a, y, bns.mean, bns.invstd, sc, t1, t2, count, dy, dres=dres,
                                              chanmul=chanmul, count_dev=cnt, fshift=fs))
        return
    P.bwd.append(lambda: ops.bn_bwd_reduce(g, a, y, bns.mean, bns.invstd, bns.dbeta, bns.dgamma, chanmul=chanmul,
                                           scale=sc, fshift=fs))
    cdev = None
    if mode == ops.BN_TRAIN_SYNC:
        cdev = st[2 * bns.C:2 * bns.C + 1]
        if P.dist is not None:
            # dgamma|dbeta are adjacent in the flat gradient buffer: one all-reduce for both
            both = P.gflat[bns.dgamma.storage_offset():bns.dgamma.storage_offset() + 2 * C]
            P.bwd.append(lambda: P.dist.all_reduce(both))
    P.bwd.append(lambda: ops.bn_bwd_apply(g, a, y, bns.mean, bns.invstd, sc, bns.dbeta, bns.dgamma, count, dy, dres=dres,
                                          chanmul=chanmul, count_dev=cdev, fshift=fs))
    if mode == ops.BN_TRAIN_SYNC and P.dist is not None:
        # the bucket all-reduce at the end sums gflat over ranks again: pre-divide the already-global dgamma/dbeta
        both = P.gflat[bns.dgamma.storage_offset():bns.dgamma.storage_offset() + 2 * C]
        P.bwd.append(lambda: both.mul_(1.0 / P.world))


class ConvBNRec:
    """Conv2d -> SynchronizedBatchNorm2d -> (+shortcut) -> ReLU -> (Dropout2d mask).
    Reference: Bottleneck/BasicBlock.forward (models/resnet.py:37-53,72-92), conv3x3_bn_relu (models/models.py:160-167),
    PPM branches / conv_last (models/models.py:443-462)."""

    def __init__(self, P, xs, cw, bns, relu, res, chanmul, apply, post_add=None):
        self.P, self.xs, self.cw, self.bns, self.relu, self.res, self.chanmul, self.apply = P, xs, cw, bns, relu, res, chanmul, apply
        self.post_add = post_add
        assert post_add is None or (res is None and relu and apply and chanmul is None)
        srcs = [x.t for x in xs]
        assert sum(s.shape[3] for s in srcs) == cw.I
        self.geom, ho, wo = P._conv_geom(srcs, cw)
        n = srcs[0].shape[0]
        assert bns.C == cw.O
        self.mode = P._bn_mode(bns)
        self.count = n * ho * wo
        self.fused = False  # set by the (single) consumer when its dgrad epilogue does this layer's BN-backward reduce
        self.coop = False   # conv + train-mode BN as one persistent kernel (_try_coop)
        self.dy_pre = None  # set when the consumer's data-gradient kernel already produced this layer's dy
        self.folded = (P.fold_bn_eval and self.mode == ops.BN_EVAL and not P.with_grad and chanmul is None and
                       (res is None or isinstance(res, Act) or res.folded))
        if self.folded:
            self._init_folded(n, ho, wo)
            return
        if self._try_coop(n, ho, wo):
            return
        self.y = P._new(n, ho, wo, bns.Cp)   # channels >= cw.O are written as zeros by the conv kernel
        st = bns.stats
        train = self.mode != ops.BN_EVAL
        C = cw.O
        geom, wf, y = self.geom, cw.wf, self.y
        P._need_weights(cw)
        m = bns.mod
        # single-GPU train mode: the LAST CTA of the conv launch finalises the statistics it has just completed
        # (sseg_conv_igemm_bnfin) - one dependent launch less per layer on the step's critical path
        self.bnfin = (P.fuse_bnfin and self.mode == ops.BN_TRAIN and bns.counter is not None and not P.fuse_finalize)
        if self.bnfin:
            upd = m.track_running_stats and m.running_mean is not None
            bn = ops.make_bn_fused(m.weight.detach() if m.weight is not None else None,
                                   m.bias.detach() if m.bias is not None else None, m.eps,
                                   m.momentum if m.momentum is not None else 0.1, self.count, st[:C], st[C:2 * C],
                                   bns.counter, bns.mean[:C], bns.invstd[:C], bns.scale[:C], bns.shift[:C],
                                   running_mean=m.running_mean if upd else None, running_var=m.running_var if upd else None)
            P.keep.append(bn)
            P.fwd.append(lambda: ops.conv_igemm_bnfin(geom, wf, C, y, bn))
        else:
            P.fwd.append(lambda: ops.conv_igemm(geom, wf, C, y, stat_sum=st[:C] if train else None,
                                                stat_sqsum=st[C:2 * C] if train else None))
        if apply:
            self.a = P._new_act(n, ho, wo, cw.O)
            self.a.producer = self
            r = rs = rb = None
            if isinstance(res, Act):
                r = res.tp
            elif isinstance(res, ConvBNRec):
                assert not res.folded
                r, rs, rb = res.y, res.bns.scale, res.bns.shift
            if post_add is not None:
                r = post_add.tp
            _emit_bn_forward(P, bns, self.mode, self.count, y, self.a.tp, relu, r, rs, rb, chanmul,
                             res_after_relu=post_add is not None, finalized=self.bnfin)
        else:
            self.a = None
            _emit_bn_forward(P, bns, self.mode, self.count, y, None, False, None, None, None, None, finalized=self.bnfin)

    def _try_coop(self, n, ho, wo):
        """Training, single-GPU BN: conv + statistics + normalise (+shortcut, ReLU, dropout mask) in one launch when the
        layer fits (sseg_conv_bn_train). Returns False to fall back to conv / finalize / apply."""
        P, cw, bns = self.P, self.cw, self.bns
        self.coop = False
        sync = self.mode == ops.BN_TRAIN_SYNC and P.peer is not None   # statistics pooled over NVLink inside the kernel
        if not (P.coop_bn and (self.mode == ops.BN_TRAIN or sync) and self.apply and bns.counter is not None):
            return False
        if self.mode == ops.BN_TRAIN and P.peer is not None:
            return False
        if isinstance(self.res, ConvBNRec) and getattr(self.res, "coop", False):
            return False
        m = bns.mod
        C = cw.O
        y = P._new(n, ho, wo, bns.Cp) if P.with_grad else None   # the backward pass reads the raw conv output
        a = P._new_act(n, ho, wo, C)
        r = rs = rb = None
        if isinstance(self.res, Act):
            r = self.res.tp
        elif isinstance(self.res, ConvBNRec):
            r, rs, rb = self.res.y, self.res.bns.scale, self.res.bns.shift
        if self.post_add is not None:
            r = self.post_add.tp
        upd = m.track_running_stats and m.running_mean is not None
        st = bns.stats
        extra = {}
        if sync:
            # [sum | sqsum | count] of this layer sit at stats_off in EVERY rank's arena; flags [flag_off, +world)
            P.sinit[bns.stats_off + 2 * C] = float(self.count)
            peer = ops.make_coop_peer(P.peer, bns.stats_off, C, bns.flag_off, P.peer_step)
            extra = dict(peer=peer, count_out=bns.tot[2 * bns.Cp:2 * bns.Cp + 1])
            if upd:
                extra.update(tmp_running_mean=m._tmp_running_mean, tmp_running_var=m._tmp_running_var,
                             running_iter=m._running_iter)
        bn = ops.make_bn_fused(m.weight.detach() if m.weight is not None else None,
                               m.bias.detach() if m.bias is not None else None, m.eps,
                               m.momentum if m.momentum is not None else 0.1, self.count, st[:C], st[C:2 * C], bns.counter,
                               bns.mean[:C], bns.invstd[:C], bns.scale[:C], bns.shift[:C],
                               running_mean=m.running_mean if (upd and not sync) else None,
                               running_var=m.running_var if (upd and not sync) else None,
                               res=r, rscale=rs, rshift=rb, chanmul=self.chanmul, relu=self.relu,
                               res_after_relu=self.post_add is not None, **extra)
        geom, wf = self.geom, cw.wf
        if P.dry_run:
            fits = _coop_fits_estimate(n, ho, wo, cw.k, C)
        else:
            fits = ops.conv_bn_train_fits(geom, wf, C, y, a.tp, bn)
        if not fits:
            return False
        self.coop, self.y, self.a = True, y, a
        self.a.producer = self
        P.keep.append(bn)
        P._need_weights(cw)
        P.fwd.append(lambda: ops.conv_bn_train(geom, wf, C, y, a.tp, bn))
        if sync and upd:   # running_mean / running_var = accumulators / running_iter (batchnorm.py:136-137)
            P.fwd.append(lambda: ops.bn_running_from_tmp(m._tmp_running_mean, m._tmp_running_var, m._running_iter,
                                                         m.running_mean, m.running_var))
        return True

    def _init_folded(self, n, ho, wo):
        """Inference: scale/shift come from the running statistics alone (finalize first), then ONE kernel computes
        relu(conv * scale + shift (+ shortcut)). A projection shortcut built this way holds its finished BN output in
        `y`, which the consuming record simply adds."""
        P, cw, bns = self.P, self.cw, self.bns
        _emit_bn_forward(P, bns, self.mode, self.count, None, None, False, None, None, None, None)
        if self.apply:
            self.a = P._new_act(n, ho, wo, cw.O)
            self.a.producer = self
            out, self.y = self.a.tp, None
        else:
            self.a = None
            out = self.y = P._new(n, ho, wo, bns.Cp)
        addend = None
        if isinstance(self.res, Act):
            addend = self.res.tp
        elif isinstance(self.res, ConvBNRec):
            addend = self.res.y
        relu = ops.RELU_AFTER_ADD if self.relu else ops.RELU_NONE
        if self.post_add is not None:
            addend, relu = self.post_add.tp, ops.RELU_BEFORE_ADD
        geom, wf, C = self.geom, cw.wf, cw.O
        P._need_weights(cw)
        P.fwd.append(lambda: ops.conv_igemm_affine(geom, wf, C, out, bns.scale, bns.shift, relu=relu, addend=addend))

    def backward(self, g_override=None):
        P, cw, bns = self.P, self.cw, self.bns
        assert not self.folded, "folded (inference) records have no backward"
        if not self.apply and g_override is None:
            return  # projection shortcuts are driven by the record that consumed them
        g = g_override if g_override is not None else self.a.g
        if g is None:
            return
        if self.dy_pre is not None:
            # sseg_conv_dgrad_bn (emitted by the consumer) did this layer's BN backward: dy, dgamma, dbeta are final
            assert self.fused and self.res is None and self.post_add is None
            dy = self.dy_pre
            geom, gw, O = self.geom, cw.gw, cw.O
            P.bwd.append(P.on_side(lambda: ops.conv_wgrad(geom, dy, O, gw)))
            self._emit_dgrad(dy)
            return
        dy = torch.empty_like(self.y)
        dres = None
        ds_rec = None
        if self.post_add is not None:
            # d(out)/d(post_add) = identity: the top-down branch reads the very same gradient tensor (no copy)
            assert self.post_add.g is None and not self.post_add.gw, "post-add input must have a single consumer"
            self.post_add.g, self.post_add.gw = g, True
        if isinstance(self.res, Act):
            dres, acc = P.grad_target(self.res)
            assert not acc, "identity shortcut must be the first gradient contribution of the block input"
        elif isinstance(self.res, ConvBNRec):
            ds_rec = self.res
            dres = torch.empty_like(ds_rec.y)
        has_relu = self.apply and self.relu
        from_y = has_relu and self.res is None          # no shortcut: the mask is a function of y alone
        a = self.a.tp if (has_relu and not from_y) else None
        _emit_bn_backward(P, bns, self.mode, self.count, g, a, self.y, dy, dres, self.chanmul, mask_from_y=from_y,
                          fused=self.fused)
        if ds_rec is not None:
            ds_rec.backward(g_override=dres)
        # weight gradient: GEMM over pixels (sseg_conv_wgrad)
        geom, gw, O = self.geom, cw.gw, cw.O
        P.bwd.append(P.on_side(lambda: ops.conv_wgrad(geom, dy, O, gw)))
        # data gradient: implicit GEMM of dy with the transposed weight, taps mirrored
        self._emit_dgrad(dy)

    def _emit_dgrad(self, dy):
        P, cw = self.P, self.cw
        xs = self.xs
        if cw.stride == 1:
            dh, dw = ops.conv_taps(cw.k, cw.dil)
            gd = ops.make_geom([dy], ([-v for v in dh], [-v for v in dw]), tap_koff=[t * cw.Opad for t in range(cw.T)])
            P.keep.append(gd)
            if len(xs) == 1:
                buf, acc = P.grad_target(xs[0])
            else:
                # virtual concat: one gradient tensor, each source's gradient is a channel slice of it
                n, h, w, _ = xs[0].t.shape
                buf = P._new(n, h, w, cw.I)
                off = 0
                for x in xs:
                    assert not x.gw, "concat sources must receive their first gradient from the concat conv"
                    c = x.t.shape[3]
                    x.g, x.gw = buf[..., off:off + c], True
                    off += c
                acc = False
            wd, I = cw.wd, cw.I
            prod = xs[0].producer if len(xs) == 1 else None
            common = (P.fuse_bnbwd and prod is not None and prod.apply and prod.relu and prod.post_add is None and
                      prod.chanmul is None and I % 8 == 0 and
                      not (prod.mode == ops.BN_TRAIN_SYNC and P.peer is None and P.dist is not None))
            fuse = common and not acc and xs[0].uses == 1 and prod.res is None
            # a producer WITH a shortcut (the output of a residual block): its gradient has several contributions (the next
            # block's first conv and its shortcut); the launch that adds the LAST one holds the complete gradient in its
            # epilogue and can reduce it there, the ReLU mask coming from the producer's saved output
            fuse_res = (common and P.fuse_bnbwd_res and not fuse and isinstance(prod, ConvBNRec) and prod.res is not None and
                        xs[0].gcount == xs[0].uses and not prod.folded)
            if fuse_res:
                prod.fused = True
                pb = prod.bns
                if prod.mode == ops.BN_TRAIN_SYNC and P.peer is not None:
                    s1, s2 = pb.part[:pb.Cp], pb.part[pb.Cp:2 * pb.Cp]
                else:
                    s1, s2 = pb.dbeta, pb.s2y
                py, pa = prod.y, xs[0].tp
                P.bwd.append(lambda: ops.conv_igemm_bnbwd_res(gd, wd, I, buf, py, pa, s1, s2, addend=buf if acc else None))
            elif fuse:
                # single consumer, no shortcut: the BN-backward reduction of the producer rides in this dgrad's epilogue
                prod.fused = True
                pb = prod.bns
                if prod.mode == ops.BN_TRAIN_SYNC and P.peer is not None:
                    s1, s2 = pb.part[:pb.Cp], pb.part[pb.Cp:2 * pb.Cp]
                else:
                    s1, s2 = pb.dbeta, pb.s2y
                py = prod.y
                sync = prod.mode == ops.BN_TRAIN_SYNC and P.peer is not None
                coop = (P.coop_bn and (sync or (prod.mode == ops.BN_TRAIN and P.peer is None)) and
                        pb.counter_bwd is not None and prod.cw.O == I)
                kw = {}
                if coop:
                    # the producer's whole BN backward rides in this data-gradient kernel: dy is written, g never is
                    dyp = torch.empty_like(py)
                    n_, h_, w_, _ = py.shape
                    args = (gd, wd, I, py, dyp, pb.scale, pb.shift, pb.mean, pb.invstd, prod.count, s1, s2, pb.dgamma,
                            pb.counter_bwd)
                    coop = _coop_fits_estimate(n_, h_, w_, cw.k, I) if P.dry_run else ops.conv_dgrad_bn(*args, query=True)
                if coop and sync:
                    # partial sums [s1 | s2raw] in this rank's arena, pooled over the ranks inside the kernel; the pooled
                    # pixel count was left in tot[2*Cp] by the forward kernel; dbeta / dgamma come out divided by world
                    kw = dict(peer=ops.make_coop_peer(P.peer, pb.part_off, pb.Cp, pb.flag_off + 8, P.peer_step),
                              count_dev=pb.tot[2 * pb.Cp:2 * pb.Cp + 1], dbeta_out=pb.dbeta)
                    P.keep.append(kw["peer"])
                if coop:
                    prod.dy_pre = dyp
                    P.bwd.append(lambda: ops.conv_dgrad_bn(*args, **kw))
                else:
                    P.bwd.append(lambda: ops.conv_igemm_bnbwd(gd, wd, I, buf, py, pb.scale, pb.shift, s1, s2))
            else:
                P.bwd.append(lambda: ops.conv_igemm(gd, wd, I, buf, n_store=_pad(I, 8), addend=buf if acc else None))
        else:
            x = xs[0]
            buf, acc = P.grad_target(x)
            dh, dw, src = ops.conv_s2_taps(cw.k)
            planes = ops.parity_planes(buf)
            P.keep.append(planes)
            for pl in range(4):
                taps = [t for t in range(cw.T) if src[t] == pl]
                if not taps:
                    if not acc:
                        P.bwd.append(lambda v=planes[pl]: v.zero_())
                    continue
                gd = ops.make_geom([dy], ([-dh[t] for t in taps], [-dw[t] for t in taps]),
                                   tap_koff=[t * cw.Opad for t in taps])
                P.keep.append(gd)
                wd, I, view = cw.wd, cw.I, planes[pl]
                P.bwd.append(lambda gd=gd, view=view: ops.conv_igemm(gd, wd, I, view, n_store=_pad(I, 8),
                                                                     addend=view if acc else None))


class SumRec:
    """One output of an HRNet exchange unit: a = ReLU(sum of terms), models/hrnet.py:225-250.
    terms: ("id", Act) the branch's own tensor | ("same", ConvBNRec apply=False) conv+BN at this resolution |
    ("up", ConvBNRec apply=False) conv+BN at a lower resolution, bilinearly sampled inside the sum kernel."""

    def __init__(self, P, terms):
        self.P, self.terms = P, terms
        ident = [t for k, t in terms if k == "id"]
        assert len(ident) == 1 and 2 <= len(terms) <= 4
        self.ident = ident[0]
        self.ident.uses += 1
        n, h, w, c = self.ident.t.shape
        self.a = Act(P._new(n, h, w, c))
        tup = []
        for kind, t in terms:
            if kind == "id":
                tup.append((t.t, None, None))
            else:
                assert t.bns.C == c and t.bns.Cp == c
                tup.append((t.y, None, None) if t.folded else (t.y, t.bns.scale, t.bns.shift))
        arr = ops.make_sum_terms(tup)
        P.keep.append((arr, tup))
        out = self.a.t
        P.fwd.append(lambda: ops.sum_terms(arr, out, relu=True))

    def backward(self):
        P = self.P
        g = self.a.g
        if g is None:
            return
        ds = torch.empty_like(self.a.t)   # gradient w.r.t. the pre-ReLU sum = gradient of every term
        idbuf, acc = P.grad_target(self.ident)
        out = self.a.t
        P.bwd.append(lambda: ops.relu_mask_bwd(g, out, ds, idbuf, accumulate=acc))
        for kind, rec in self.terms:
            if kind == "same":
                rec.backward(g_override=ds)
            elif kind == "up":
                glow = torch.empty_like(rec.y)   # adjoint of the bilinear sampling, at the term's own resolution
                n, ho, _, c = ds.shape
                scratch = P._new(n * ho * glow.shape[2] * c, dtype=torch.float32)
                P.bwd.append(lambda glow=glow, scratch=scratch: ops.bilinear_bwd(ds, glow, accumulate=False,
                                                                                 scratch=scratch))
                rec.backward(g_override=glow)


class MaxPoolRec:
    """nn.MaxPool2d(3, 2, 1): models/resnet.py:109."""

    def __init__(self, P, x):
        self.P, self.x = P, x
        x.uses += 1
        n, h, w, c = x.t.shape
        ho, wo = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
        self.a = Act(P._new(n, ho, wo, c))
        self.idx = P._new(n, ho, wo, c, dtype=torch.uint8) if P.with_grad else None
        P.fwd.append(lambda: ops.maxpool_fwd(x.t, self.a.t, self.idx))

    def backward(self):
        P = self.P
        if self.a.g is None:
            return
        buf, acc = P.grad_target(self.x)
        assert not acc
        g = self.a.g
        assert g.is_contiguous()
        P.bwd.append(lambda: ops.maxpool_bwd(g, self.idx, buf))


class AvgPoolRec:
    """nn.AdaptiveAvgPool2d(scale): models/models.py:447.  The pools of one pyramid share their input; their backward
    passes are emitted as ONE kernel (sseg_avgpool_bwd takes up to 4 scales) once the last of them is reached."""

    def __init__(self, P, x, scale):
        self.P, self.x, self.scale = P, x, scale
        x.uses += 1
        n, h, w, c = x.t.shape
        self.a = Act(P._new(n, scale, scale, c))
        P.fwd.append(lambda: ops.avgpool_fwd(x.t, scale, self.a.t))
        self.group = P.pool_groups.setdefault(id(x), [])
        self.group.append(self)
        self.done = False

    def backward(self):
        P = self.P
        self.done = True
        if not all(r.done for r in self.group):
            return  # an earlier-in-forward sibling comes later in backward and emits the fused kernel
        live = [r for r in self.group if r.a.g is not None]
        if not live:
            return
        buf, acc = P.grad_target(self.x)
        gs, ss = [r.a.g for r in live], [r.scale for r in live]
        for i in range(0, len(live), 4):
            first = i == 0
            P.bwd.append(lambda gs=gs[i:i + 4], ss=ss[i:i + 4], base=(buf if (acc or not first) else None):
                         ops.avgpool_bwd(base, gs, ss, buf))


class UpsampleRec:
    """F.interpolate(bilinear, align_corners=False) of a PPM branch to the conv5 size: models/models.py:472-475."""

    def __init__(self, P, x, h, w):
        self.P, self.x = P, x
        x.uses += 1
        n, _, _, c = x.t.shape
        self.a = Act(P._new(n, h, w, c))
        P.fwd.append(lambda: ops.bilinear_fwd(x.t, self.a.t))

    def backward(self):
        P = self.P
        if self.a.g is None:
            return
        buf, acc = P.grad_target(self.x)
        g = self.a.g
        n, ho, _, c = g.shape
        scratch = P._new(n * ho * buf.shape[2] * c, dtype=torch.float32)
        P.bwd.append(lambda: ops.bilinear_bwd(g, buf, accumulate=acc, scratch=scratch))


class ClassifierRec:
    """The num_class 1x1 conv with bias (conv_last[4] / conv_last_deepsup): models/models.py:462,465."""

    def __init__(self, P, x, cw):
        self.P, self.x, self.cw = P, x, cw
        x.uses += 1
        n, h, w, _ = x.t.shape
        self.ld = _pad(cw.O, 8) + 8
        self.logits = P._new(n, h, w, self.ld, dtype=torch.float32, zero=True)
        self.geom, _, _ = P._conv_geom([x.t], cw)
        self.dlogits = None
        geom, wf, O, bias = self.geom, cw.wf, cw.O, cw.mod.bias
        P._need_weights(cw)
        P.fwd.append(lambda: ops.conv_igemm(geom, wf, O, self.logits, n_store=_pad(O, 8),
                                            bias=bias.detach() if bias is not None else None))

    def backward(self):
        P, cw = self.P, self.cw
        if self.dlogits is None:
            return
        dl = self.dlogits  # bf16 [n,h,w,Opad], zero padded
        geom, gw, O = self.geom, cw.gw, cw.O
        gb = cw.gb

        def head_wgrad():
            ops.conv_wgrad(geom, dl, O, gw)
            if gb is not None:
                ops.colsum(dl, O, gb)
        P.bwd.append(P.on_side(head_wgrad))
        buf, acc = P.grad_target(self.x)
        gd = ops.make_geom([dl], ([0], [0]), tap_koff=[0])
        P.keep.append(gd)
        wd, I = cw.wd, cw.I
        P.bwd.append(lambda: ops.conv_igemm(gd, wd, I, buf, n_store=_pad(I, 8), addend=buf if acc else None))


class LossRec:
    """log_softmax + NLLLoss(ignore_index=-1) (+ deep supervision) + pixel_acc: models/models.py:12-18,37-42,492-493."""

    def __init__(self, P):
        self.P = P
        self.heads = [r for r in P.records if isinstance(r, ClassifierRec)]
        n, h, w, _ = P.logits.shape
        self.lse = [P._new(n * h * w, dtype=torch.float32) for _ in self.heads]
        C = P.num_class
        accs = [P.acc_main, P.acc_ds]
        for head, lse, acc in zip(self.heads, self.lse, accs):
            lg = head.logits[..., :_pad(C, 8)]
            P.fwd.append(lambda lg=lg, lse=lse, acc=acc: ops.softmax_nll_fwd(lg, C, P.label, lse, acc))
        ds = P.seg.deep_sup_scale if len(self.heads) > 1 else None
        self.weights = [1.0] + ([float(ds)] if ds is not None else [])
        P.fwd.append(lambda: ops.nll_finalize(P.acc_main, P.acc_ds if ds is not None else None, ds or 0.0, P.out))

    def backward(self):
        P = self.P
        C = P.num_class
        accs = [P.acc_main, P.acc_ds]
        for head, lse, acc, wgt in zip(self.heads, self.lse, accs, self.weights):
            n, h, w, _ = head.logits.shape
            head.dlogits = P._new(n, h, w, head.cw.Opad)
            lg, dl = head.logits[..., :_pad(C, 8)], head.dlogits
            wgt = wgt / P.world   # mean over the ranks' losses (train.py:42): every gradient below inherits the factor
            P.bwd.append(lambda lg=lg, lse=lse, acc=acc, wgt=wgt, dl=dl: ops.softmax_nll_bwd(lg, C, P.label, lse, acc,
                                                                                            wgt, dl))
