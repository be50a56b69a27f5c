"""Entry points the nn.Module classes call: they look up (or build) the step program for the input shape, run it, and
connect its precomputed gradients to autograd.  See engine/program.py for what a program is."""
import collections
import os

import torch

from .program import SegProgram

_MAX_PROGRAMS = 8            # per module; inference over many image sizes would otherwise pin GBs of activations
_MAX_PROGRAM_BYTES = int(float(os.environ.get("SSEG_PROGRAM_CACHE_GB", "48")) * 2 ** 30)   # ... and so would many crops
_CAPTURE_ON_USE = int(os.environ.get("SSEG_CAPTURE_ON_USE", "3"))   # a shape is graph-captured on its n-th use


def _require_cuda_module(mod):
    p = next(mod.parameters(), None)
    if p is None or not p.is_cuda:
        raise RuntimeError("the B200 engine has no CPU path: move the module to a CUDA device (inputs may stay on the "
                           "host - they are copied into the step's static device buffers)")


def _prog_bytes(prog):
    """Device bytes a program pins (activations, gradients, operand copies), from the allocator's own bookkeeping."""
    return getattr(prog, "_cache_bytes", 0)


def _evict(cache):
    while len(cache) > 1 and (len(cache) > _MAX_PROGRAMS or sum(_prog_bytes(q) for q in cache.values()) > _MAX_PROGRAM_BYTES):
        cache.popitem(last=False)


def _programs(mod):
    cache = mod.__dict__.get("_b200_programs")
    if cache is None:
        cache = collections.OrderedDict()
        mod.__dict__["_b200_programs"] = cache  # not a Module attribute: invisible to state_dict / replicate
    return cache


_TREE_RECHECK = 256   # calls between two full re-walks of a module tree (see _tree)


def _tree(mod):
    """(modules, parameters) of `mod` as flat lists, cached on the module. nn.Module.modules() / .parameters() re-walk the
    whole tree through nested generators on every call (~0.1-0.3 ms each for ResNet50 + PPM: most of the host time of a
    step, during which the GPU sits idle between `loss.item()` and the next graph launch). The cache is dropped when a
    direct child was replaced, and re-validated against a full walk every _TREE_RECHECK calls: restructuring a model
    between two steps is picked up then, together with its step programs."""
    kids = tuple(id(m) for m in mod._modules.values())
    c = mod.__dict__.get("_b200_tree")
    if c is not None and c[0] == kids and c[3] > 0:
        c[3] -= 1
        return c[1], c[2]
    modules, params = list(mod.modules()), list(mod.parameters())
    if c is not None and ([id(m) for m in c[1]] != [id(m) for m in modules] or [id(q) for q in c[2]] != [id(q) for q in params]):
        mod.__dict__.pop("_b200_programs", None)   # compiled from a tree that no longer exists
    mod.__dict__["_b200_tree"] = [kids, modules, params, _TREE_RECHECK]
    return modules, params


def fast_zero_grad(mod, set_to_none=True):
    """nn.Module.zero_grad over the cached parameter list (same semantics)."""
    for p in _tree(mod)[1]:
        if p.grad is not None:
            if set_to_none:
                p.grad = None
            else:
                if p.grad.grad_fn is not None:
                    p.grad.detach_()
                else:
                    p.grad.requires_grad_(False)
                p.grad.zero_()


def get_program(seg, img_shape, seg_size=None, with_grad=None, dropout_masks=None, capture=None, inputs=None,
                head_out=None, head_weight=1.0):
    modules, params = _tree(seg)
    if with_grad is None:
        with_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    bn_flags = tuple([m.training for m in modules])
    head = None if head_out is None else (head_out.data_ptr(), float(head_weight))
    key = (tuple(img_shape), seg_size, bool(with_grad), hash(bn_flags), id(dropout_masks), head)
    cache = _programs(seg)
    prog = cache.get(key)
    if prog is None:
        dev = params[0].device
        before = torch.cuda.memory_allocated(dev) if dev.type == "cuda" and torch.cuda.is_available() else 0
        prog = SegProgram(seg, tuple(img_shape), training=seg.training, with_grad=with_grad, seg_size=seg_size,
                          dropout_masks=dropout_masks, head_out=head_out, head_weight=head_weight)
        after = torch.cuda.memory_allocated(dev) if dev.type == "cuda" and torch.cuda.is_available() else 0
        prog._cache_bytes = max(0, after - before)
        if inputs is not None:
            prog.load_inputs(*inputs)  # the capture warm-up runs the step: give it real data, not uninitialised memory
        if capture:
            prog.capture()  # callers that know the shape recurs (bench, tests) capture at once
        cache[key] = prog
        _evict(cache)
    else:
        cache.move_to_end(key)
    return prog


def _maybe_capture(prog):
    """The reference loader draws a new batch shape almost every iteration (5 short sizes x free aspect ratio, padded to
    8: hundreds of shapes), so a program runs eagerly until its shape has recurred `SSEG_CAPTURE_ON_USE` times; only then
    is it captured into a CUDA graph (capture = two warm-up executions + the capture pass). Under torch.distributed the
    ranks draw DIFFERENT shapes: capture never issues a collective (SegProgram.capture stubs them during its warm-up), so
    ranks need not capture in lockstep."""
    prog.uses = getattr(prog, "uses", 0) + 1
    if prog.graph is None and prog.uses == _CAPTURE_ON_USE and not getattr(prog, "no_capture", False):
        prog.capture(warm=_CAPTURE_ON_USE <= 1)   # it has run eagerly before: nothing to warm up, nothing executed here


class _TrainStep(torch.autograd.Function):
    """forward = the whole fwd+bwd program (one CUDA-graph replay); backward = hand the gradients to autograd."""

    @staticmethod
    def forward(ctx, prog, *params):
        prog.run()
        prog.generation = getattr(prog, "generation", 0) + 1
        ctx.prog, ctx.params, ctx.generation, ctx.consumed = prog, params, prog.generation, False
        out = prog.out.clone()
        loss, acc = out[0], out[1]
        ctx.mark_non_differentiable(acc)
        return loss, acc

    @staticmethod
    def backward(ctx, g_loss, g_acc):
        prog = ctx.prog
        if not prog.with_grad:
            raise RuntimeError("this program was built without a backward schedule")
        if ctx.consumed or prog.generation != ctx.generation:
            raise RuntimeError("the step's gradients live in the program's static buffers and were already consumed or "
                               "overwritten by a later forward pass of the same shape: call backward() once, right after "
                               "the forward pass it belongs to")
        ctx.consumed = True
        # chain rule: every gradient of the step times the incoming d(loss). The buffers are scaled in place by a DEVICE
        # scalar (no host sync); `loss.backward()` seeds exactly 1, for which the kernel exits without touching memory.
        from . import ops
        g = g_loss.detach().to(torch.float32).reshape(1).contiguous()
        ops.scale_by_scalar(prog.gflat, g)   # every parameter gradient is a view of this one buffer
        grads = prog.param_grads()
        # fresh views: autograd then adopts the static buffers as .grad instead of cloning 200 MB per step
        return (None,) + tuple(_alias(grads[p]) if p in grads else None for p in ctx.params)


def _alias(t):
    """A new tensor object over the same memory and strides (autograd adopts it as .grad without a copy)."""
    return t.as_strided(t.shape, t.stride())


def _unalias_grads(prog, params):
    """Gradient accumulation (backward() of several steps without zero_grad() in between): a .grad that still aliases this
    program's static buffer would be overwritten by the run that is about to start - give it storage of its own."""
    grads = getattr(prog, "_grad_ptrs", None)
    if grads is None:
        pg = prog.param_grads()
        grads = prog._grad_ptrs = {id(p): pg[p].data_ptr() for p in params if p in pg}
    for p in params:
        if p.grad is not None and grads.get(id(p)) == p.grad.data_ptr():
            p.grad = p.grad.clone()


def segmentation_train_step(seg, img, label):
    """SegmentationModule.forward, training branch (reference models/models.py:31-43) -> (loss, acc).
    `img` / `label` may live on the host (what the reference's loader + `--gpus 0` hand over, train.py:176-183): they are
    copied straight into the step's static device buffers (asynchronously when pinned)."""
    _require_cuda_module(seg)
    prog = get_program(seg, img.shape, inputs=(img, label))
    prog.load_inputs(img, label)
    _maybe_capture(prog)
    if prog.with_grad:
        params = [p for p in _tree(seg)[1] if p.requires_grad]
        _unalias_grads(prog, params)
        return _TrainStep.apply(prog, *params)
    prog.run()
    out = prog.out.clone()
    return out[0], out[1]


def segmentation_inference(seg, img, seg_size):
    """SegmentationModule.forward, inference branch (reference models/models.py:44-47): softmax probabilities of the
    main head, bilinearly up-sampled to seg_size, fp32 NCHW."""
    _require_cuda_module(seg)
    if not getattr(seg.decoder, "use_softmax", False):
        raise RuntimeError("inference (segSize=...) requires a decoder built with use_softmax=True")
    if os.environ.get("SSEG_ACCURATE_INFERENCE", "0") == "1":
        # fp32-accurate mode (bf16 pairs, three-term products; engine/accurate.py): logits within 1e-3 of the fp32 reference
        from .accurate import AccurateInference
        key = ("acc", tuple(img.shape), tuple(seg_size), _flags(seg))
        prog = _cached(seg, key, lambda: AccurateInference(seg, tuple(img.shape), tuple(seg_size)))
    else:
        prog = get_program(seg, img.shape, seg_size=tuple(seg_size), with_grad=False, capture=False)
    prog.load_inputs(img)
    # image sizes vary during evaluation, so a program is only captured into a CUDA graph once its shape recurs
    prog.uses = getattr(prog, "uses", 0) + 1
    if prog.uses == 2 and prog.graph is None:
        prog.capture()
    prog.run()
    return prog.probs.clone()


def shard_scales(num_scales, world, rank):
    """Scale k runs on rank k mod world: the scales are independent units."""
    return [k for k in range(num_scales) if k % world == rank]


def multiscale_inference(seg, imgs, seg_size, group=None, _run_scale=None):
    """eval.py:58-75 — `scores = sum_k segmentation_module({img_k}, segSize) / len(imgs)` for the resized copies `imgs`
    (list of [N,3,h_k,w_k] tensors) of one image — without the three full passes over the 150-channel score map the
    reference spends per scale: every scale's head kernel adds `softmax / len(imgs)` straight into ONE score map.

    With torch.distributed initialised (world size G) the scales are sharded k -> rank k mod G and the partial score
    maps summed with one all-reduce, so every rank returns the full [N,C,*seg_size] map (argmax it like eval.py:74).
    `_run_scale(img, scores, weight)` replaces the engine call in the CPU tests of this host logic."""
    import torch.distributed as dist
    seg_size = tuple(seg_size)
    sharded = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    world, rank = (dist.get_world_size(group), dist.get_rank(group)) if sharded else (1, 0)
    if _run_scale is None:
        _require_cuda_module(seg)
        if not getattr(seg.decoder, "use_softmax", False):
            raise RuntimeError("inference (segSize=...) requires a decoder built with use_softmax=True")
        num_class = [m for m in seg.decoder.modules() if isinstance(m, torch.nn.Conv2d)][-1].out_channels
    else:
        num_class = seg.num_class
    n = imgs[0].shape[0]
    bufs = seg.__dict__.setdefault("_b200_scores", {})
    shape = (n, num_class) + seg_size
    scores = bufs.get((shape, imgs[0].device))
    if scores is None:
        if len(bufs) >= 4:
            bufs.clear()   # evaluation images come in many sizes: do not pin a score map per size forever
        scores = bufs[(shape, imgs[0].device)] = torch.empty(shape, device=imgs[0].device, dtype=torch.float32)
    scores.zero_()
    weight = 1.0 / len(imgs)
    for k in shard_scales(len(imgs), world, rank):
        if _run_scale is not None:
            _run_scale(imgs[k], scores, weight)
            continue
        prog = get_program(seg, imgs[k].shape, seg_size=seg_size, with_grad=False, capture=False, head_out=scores,
                           head_weight=weight)
        prog.load_inputs(imgs[k])
        prog.uses = getattr(prog, "uses", 0) + 1
        if prog.uses == 2 and prog.graph is None:
            saved = scores.clone()   # the capture warm-up executes the (accumulating) step: undo its contribution
            prog.capture()
            scores.copy_(saved)
        prog.run()
    if sharded:
        dist.all_reduce(scores, group=group)
    return scores.clone()


def _cached(mod, key, build):
    cache = _programs(mod)
    prog = cache.get(key)
    if prog is None:
        prog = build()
        cache[key] = prog
        _evict(cache)
    else:
        cache.move_to_end(key)
    return prog


def _flags(mod):
    return hash(tuple([m.training for m in _tree(mod)[0]]))


def encoder_forward(enc, x):
    """Resnet / ResnetDilated called on their own (reference models/models.py:190-205,253-268): the four stage outputs
    as fp32 NCHW tensors. Forward only: gradients flow through the fused SegmentationModule program, not through
    module-level calls."""
    _require_cuda_module(enc)
    prog = _cached(enc, ("enc", tuple(x.shape), _flags(enc)),
                   lambda: SegProgram(None, tuple(x.shape), training=enc.training, with_grad=False, part="encoder", enc=enc))
    prog.load_inputs(x)
    prog.run()
    return [t.clone() for t in prog.feat_out]


def decoder_forward(dec, conv_out, seg_size):
    """PPM / PPMDeepsup / C1 / C1DeepSup called on their own (reference models/models.py:339-385,408-495) with fp32
    NCHW feature maps: log-probabilities at feature resolution (a tuple with the deep-supervision head for *_deepsup),
    or, for use_softmax decoders, probabilities up-sampled to seg_size. Forward only."""
    _require_cuda_module(dec)
    infer = bool(getattr(dec, "use_softmax", False))
    if infer and seg_size is None:
        raise RuntimeError("a use_softmax decoder needs segSize")
    shapes = tuple(tuple(t.shape) for t in conv_out)
    key = ("dec", shapes, tuple(seg_size) if infer else None, _flags(dec))
    prog = _cached(dec, key, lambda: SegProgram(None, None, training=dec.training, with_grad=False,
                                                seg_size=tuple(seg_size) if infer else None, part="decoder", dec=dec,
                                                feat_shapes=shapes))
    prog.load_features(conv_out)
    prog.run()
    outs = [t.clone() for t in prog.outputs]
    return outs[0] if len(outs) == 1 else tuple(outs)


class _ModuleBatchNorm(torch.autograd.Function):
    """SynchronizedBatchNorm{1,2,3}d called on its own (reference lib/nn/modules/batchnorm.py:56-86), on the engine's BN
    kernels: NCHW fp32 -> NHWC bf16, per-channel [sum | sum of squares] (the backward-reduce kernel with g = y = x,
    mean 0, inv_std 1), - when synchronised - ONE all-reduce of the [sum | sqsum | count] message (batchnorm.py:68-70),
    finalize in the module's formula (F.batch_norm / pooled clamp(var, eps) with accumulator running statistics / eval),
    apply, back to NCHW fp32. backward: [sum g | sum g*xhat] (+ all-reduce), then the input gradient."""

    @staticmethod
    def forward(ctx, x, weight, bias, bn):
        from . import ops
        dev, C = x.device, bn.num_features
        Cp = (C + 7) // 8 * 8          # the kernels work on 8-channel groups: pad channels hold zeros throughout
        shp = x.shape
        x4 = x.reshape(shp[0], C, -1, 1).contiguous().float()
        n, _, h, w = x4.shape
        xb = torch.zeros(n, h, w, Cp, device=dev, dtype=torch.bfloat16)
        ops.nchw_f32_to_nhwc_bf16(x4, xb)
        vec = torch.zeros(4, Cp, device=dev, dtype=torch.float32)   # mean | inv_std | scale | shift
        sync = bn.is_synchronized()
        mode = ops.BN_EVAL if not bn.training else (ops.BN_TRAIN_SYNC if sync else ops.BN_TRAIN)
        count = float(n * h * w)
        msg = torch.zeros(2 * Cp + 1, device=dev, dtype=torch.float32)   # [sum | sum of squares | count]
        cdev = None
        if mode != ops.BN_EVAL:
            zeros, ones = torch.zeros(Cp, device=dev), torch.ones(Cp, device=dev)
            ops.bn_bwd_reduce(xb, None, xb, zeros, ones, msg[:Cp], msg[Cp:2 * Cp])
            if sync:
                msg[2 * Cp] = count
                dist = _dist_or_none()
                if dist is not None:
                    dist.all_reduce(msg)
                cdev = msg[2 * Cp:]
        running = (bn.running_mean, bn.running_var, getattr(bn, "_tmp_running_mean", None),
                   getattr(bn, "_tmp_running_var", None), getattr(bn, "_running_iter", None))
        upd = mode != ops.BN_EVAL and bn.track_running_stats and bn.running_mean is not None
        ops.bn_finalize(msg[:C], msg[Cp:Cp + C], count, weight, bias, bn.eps,
                        bn.momentum if bn.momentum is not None else 0.1, mode, vec[0, :C], vec[1, :C], vec[2, :C],
                        vec[3, :C], running=running, update_running=upd, count_dev=cdev)
        ob = torch.empty_like(xb)
        ops.bn_apply(xb, vec[2], vec[3], ob, relu=False)
        out = torch.empty_like(x4)
        ops.nhwc_bf16_to_nchw_f32(ob[..., :C], out)
        ctx.save_for_backward(xb, vec, msg)
        ctx.mode, ctx.count, ctx.shape, ctx.has_affine, ctx.C = mode, count, shp, weight is not None, C
        return out.reshape(shp).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        from . import ops
        xb, vec, msg = ctx.saved_tensors
        n, h, w, Cp = xb.shape
        C, dev = ctx.C, xb.device
        g4 = g.reshape(n, C, h, w).contiguous().float()
        gb = torch.zeros_like(xb)
        ops.nchw_f32_to_nhwc_bf16(g4, gb)
        s = torch.zeros(2, Cp, device=dev, dtype=torch.float32)     # dbeta = sum g | dgamma = sum g * xhat
        ops.bn_bwd_reduce(gb, None, xb, vec[0], vec[1], s[0], s[1])
        cdev, world = None, 1
        if ctx.mode == ops.BN_TRAIN_SYNC:
            dist = _dist_or_none()
            if dist is not None:
                dist.all_reduce(s)                                   # the backward message is pooled the same way
                world = dist.get_world_size()
            cdev = msg[2 * Cp:]
        dyb = torch.empty_like(xb)
        ops.bn_bwd_apply(gb, None, xb, vec[0], vec[1], vec[2], s[0], s[1], ctx.count, dyb, eval_mode=(ctx.mode == ops.BN_EVAL),
                         count_dev=cdev)
        dx = torch.empty(n, C, h, w, device=dev, dtype=torch.float32)
        ops.nhwc_bf16_to_nchw_f32(dyb[..., :C], dx)
        dw = db = None
        if ctx.has_affine:   # every rank holds the pooled sums: its share of the (later summed) parameter gradient is 1 / world
            dw, db = s[1, :C] / world, s[0, :C] / world
        return dx.reshape(ctx.shape).to(g.dtype), dw, db, None


def _dist_or_none():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


def batch_norm(bn, x):
    """SynchronizedBatchNorm{1,2,3}d.forward outside a SegmentationModule program."""
    _require_cuda_module(bn)
    if not x.is_cuda:
        x = x.to(next(bn.parameters()).device)
    return _ModuleBatchNorm.apply(x, bn.weight, bn.bias, bn)
