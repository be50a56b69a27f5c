"""FusedSGD: torch.optim.SGD(momentum, weight_decay) semantics, one kernel launch per step for all parameters
(csrc/optim.cu).  A drop-in torch.optim.Optimizer: param groups carry lr / momentum / weight_decay like torch's, so
train.py's `adjust_learning_rate` (which rewrites group['lr']) keeps working.

Gradients are read from `p.grad` at the first step and the chunk table is cached: `p.grad` must then keep pointing to
the same storage every step (true for the B200 engine, whose gradients are static program buffers); if a `.grad`
tensor is replaced the table is rebuilt."""
import torch

from . import _C

def _dense(t):
    return t.is_contiguous() or (t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last))


def _same_layout(a, b):
    """Equal strides on every dimension that has more than one element (strides of size-1 dimensions are arbitrary)."""
    return a.shape == b.shape and all(sa == sb for n, sa, sb in zip(a.shape, a.stride(), b.stride()) if n > 1)


_CHUNK = 16384   # elements per block: 4 iterations of 4 float4 triples in flight per thread


class FusedSGD(torch.optim.Optimizer):
    def __init__(self, params, lr=0.02, momentum=0.9, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay))
        self._tables = {}
        self._steps = 0

    def _build(self, gi, group):
        chunks, keep, sig = [], [], []
        for p in group["params"]:
            if p.grad is None:
                continue
            # element i of param / grad / momentum must be the same logical element: all three dense with equal strides
            # (OIHW-contiguous, or channels-last for the engine's convolution masters and their gradient views)
            assert p.is_cuda and p.dtype == torch.float32 and _same_layout(p, p.grad) and _dense(p), \
                "FusedSGD needs dense parameters whose .grad has the same memory layout"
            st = self.state[p]
            if "momentum_buffer" not in st:
                st["momentum_buffer"] = torch.zeros_like(p)   # preserve_format: same strides as the parameter
            buf = st["momentum_buffer"]
            assert _same_layout(buf, p)
            n = p.numel()
            sig.append((p.data_ptr(), p.grad.data_ptr()))
            for off in range(0, n, _CHUNK):
                m = min(_CHUNK, n - off)
                ptrs = (p.data_ptr() + 4 * off, p.grad.data_ptr() + 4 * off, buf.data_ptr() + 4 * off)
                vec4 = int(m % 4 == 0 and all(q % 16 == 0 for q in ptrs))
                chunks.append(_C.SgdChunk(ptrs[0], ptrs[1], ptrs[2], float(group["weight_decay"]), m, vec4, 0))
            keep.append((p, buf))
        if not chunks:
            return None
        arr = (_C.SgdChunk * len(chunks))(*chunks)
        dev = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(keep[0][0].device)
        return dict(dev=dev, n=len(chunks), sig=sig, keep=keep, wd=group["weight_decay"])

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        stream = _C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for gi, group in enumerate(self.param_groups):
            tab = self._tables.get(gi)
            sig = [(p.data_ptr(), p.grad.data_ptr()) for p in group["params"] if p.grad is not None]
            if tab is None or tab["sig"] != sig or tab["wd"] != group["weight_decay"]:
                tab = self._build(gi, group)
                self._tables[gi] = tab
            if tab is None:
                continue
            _C.check(_C.lib().sseg_sgd_step(_C.ptr(tab["dev"]), tab["n"], float(group["lr"]), float(group["momentum"]),
                                            0, stream))
        self._steps += 1
        return loss
