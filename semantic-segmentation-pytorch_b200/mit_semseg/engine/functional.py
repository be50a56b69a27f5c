"""Entry points the nn.Module classes call: they look up (or build) the step program for the input shape, run it, and
connect its precomputed gradients to autograd.  See engine/program.py for what a program is."""
import collections
import os

import torch

from .program import SegProgram

_MAX_PROGRAMS = 8  # per module; inference over many image sizes would otherwise pin GBs of activations


def _programs(mod):
    cache = mod.__dict__.get("_b200_programs")
    if cache is None:
        cache = collections.OrderedDict()
        mod.__dict__["_b200_programs"] = cache  # not a Module attribute: invisible to state_dict / replicate
    return cache


def get_program(seg, img_shape, seg_size=None, with_grad=None, dropout_masks=None, capture=None, inputs=None,
                head_out=None, head_weight=1.0):
    if with_grad is None:
        with_grad = torch.is_grad_enabled() and any(p.requires_grad for p in seg.parameters())
    bn_flags = tuple(m.training for m in seg.modules())
    head = None if head_out is None else (head_out.data_ptr(), float(head_weight))
    key = (tuple(img_shape), seg_size, bool(with_grad), hash(bn_flags), id(dropout_masks), head)
    cache = _programs(seg)
    prog = cache.get(key)
    if prog is None:
        prog = SegProgram(seg, tuple(img_shape), training=seg.training, with_grad=with_grad, seg_size=seg_size,
                          dropout_masks=dropout_masks, head_out=head_out, head_weight=head_weight)
        if inputs is not None:
            prog.load_inputs(*inputs)  # the capture warm-up runs the step: give it real data, not uninitialised memory
        if capture if capture is not None else (seg_size is None):
            prog.capture()  # fixed-shape training steps are replayed as one CUDA graph
        cache[key] = prog
        while len(cache) > _MAX_PROGRAMS:
            cache.popitem(last=False)
    else:
        cache.move_to_end(key)
    return prog


class _TrainStep(torch.autograd.Function):
    """forward = the whole fwd+bwd program (one CUDA-graph replay); backward = hand the gradients to autograd."""

    @staticmethod
    def forward(ctx, prog, *params):
        prog.run()
        ctx.prog, ctx.params = prog, params
        out = prog.out.clone()
        loss, acc = out[0], out[1]
        ctx.mark_non_differentiable(acc)
        return loss, acc

    @staticmethod
    def backward(ctx, g_loss, g_acc):
        prog = ctx.prog
        if not prog.with_grad:
            raise RuntimeError("this program was built without a backward schedule")
        grads = prog.param_grads()
        gl = [grads[p] for p in ctx.params if p in grads]
        scaled = torch._foreach_mul(gl, g_loss.to(torch.float32))
        it = iter(scaled)
        return (None,) + tuple(next(it) if p in grads else None for p in ctx.params)


def segmentation_train_step(seg, img, label):
    """SegmentationModule.forward, training branch (reference models/models.py:31-43) -> (loss, acc)."""
    if not img.is_cuda:
        raise RuntimeError("the B200 engine has no CPU path: move the module and the batch to a CUDA device")
    prog = get_program(seg, img.shape, inputs=(img, label))
    prog.load_inputs(img, label)
    if prog.with_grad:
        params = [p for p in seg.parameters() if p.requires_grad]
        return _TrainStep.apply(prog, *params)
    prog.run()
    out = prog.out.clone()
    return out[0], out[1]


def segmentation_inference(seg, img, seg_size):
    """SegmentationModule.forward, inference branch (reference models/models.py:44-47): softmax probabilities of the
    main head, bilinearly up-sampled to seg_size, fp32 NCHW."""
    if not img.is_cuda:
        raise RuntimeError("the B200 engine has no CPU path: move the module and the batch to a CUDA device")
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
        if not imgs[0].is_cuda:
            raise RuntimeError("the B200 engine has no CPU path: move the module and the images to a CUDA device")
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
        while len(cache) > _MAX_PROGRAMS:
            cache.popitem(last=False)
    else:
        cache.move_to_end(key)
    return prog


def _flags(mod):
    return hash(tuple(m.training for m in mod.modules()))


def encoder_forward(enc, x):
    """Resnet / ResnetDilated called on their own (reference models/models.py:190-205,253-268): the four stage outputs
    as fp32 NCHW tensors. Forward only: gradients flow through the fused SegmentationModule program, not through
    module-level calls."""
    if not x.is_cuda:
        raise RuntimeError("the B200 engine has no CPU path: move the module and the batch to a CUDA device")
    prog = _cached(enc, ("enc", tuple(x.shape), _flags(enc)),
                   lambda: SegProgram(None, tuple(x.shape), training=enc.training, with_grad=False, part="encoder", enc=enc))
    prog.load_inputs(x)
    prog.run()
    return [t.clone() for t in prog.feat_out]


def decoder_forward(dec, conv_out, seg_size):
    """PPM / PPMDeepsup / C1 / C1DeepSup called on their own (reference models/models.py:339-385,408-495) with fp32
    NCHW feature maps: log-probabilities at feature resolution (a tuple with the deep-supervision head for *_deepsup),
    or, for use_softmax decoders, probabilities up-sampled to seg_size. Forward only."""
    if not conv_out[-1].is_cuda:
        raise RuntimeError("the B200 engine has no CPU path: move the module and the batch to a CUDA device")
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


def run_block(block, x):
    raise NotImplementedError("residual blocks run inside a SegmentationModule program on the B200 engine")


def batch_norm(bn, x):
    raise NotImplementedError("SynchronizedBatchNorm runs inside a SegmentationModule program on the B200 engine")
