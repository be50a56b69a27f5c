"""Data-parallel wrappers with the reference's names.

Reference: mit_semseg/lib/nn/parallel/data_parallel.py:13-112 and lib/nn/modules/replicate.py:27-94.

The reference runs ONE process that replicates the module onto every GPU each iteration (param broadcast), runs one
Python thread per GPU and reduces gradients back to GPU 0.  The B200 design is one process per GPU
(torch.distributed / NCCL over NVLink): replicas are persistent, SyncBN statistics and the gradient bucket are
all-reduced by the engine.  `UserScatteredDataParallel` therefore:

  * stays an `nn.DataParallel` subclass (train.py wraps the model in it and replicate.py:84 asserts the type);
  * `scatter` keeps the reference contract — the input is a list with one dict per GPU — and picks this rank's
    entry (`rank` = torch.distributed rank, 0 without a process group), copying it to the device asynchronously;
  * `forward` runs the wrapped module on this process's device and returns outputs with a leading dimension
    (0-d -> 1-d, reference data_parallel.py:33-36), so `loss.mean()` / `acc.mean()` in train.py keep working.
"""
import collections.abc

import torch
import torch.cuda as cuda
import torch.nn as nn
from torch.nn.parallel import DataParallel

__all__ = ['UserScatteredDataParallel', 'user_scattered_collate', 'async_copy_to', 'DataParallelWithCallback',
           'patch_replication_callback', 'ensure_process_group']


def async_copy_to(obj, dev, main_stream=None):
    """Recursively copy tensors in dict/list structures to `dev` without blocking; non-tensors pass through."""
    if torch.is_tensor(obj):
        v = obj.cuda(dev, non_blocking=True)
        if main_stream is not None:
            v.record_stream(main_stream)
        return v
    if isinstance(obj, collections.abc.Mapping):
        return {k: async_copy_to(o, dev, main_stream) for k, o in obj.items()}
    if isinstance(obj, collections.abc.Sequence) and not isinstance(obj, (str, bytes)):
        return [async_copy_to(o, dev, main_stream) for o in obj]
    return obj


def user_scattered_collate(batch):
    return batch


def _rank_world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def ensure_process_group():
    """The reference's scripts know nothing about torch.distributed (they replicate inside one process). Launched as one
    process per GPU - `python -m torch.distributed.run --nproc-per-node N train.py --gpus 0-(N-1)` - the launcher's
    environment (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*) is all there is, so the first multi-GPU call the script makes
    (wrapping the model in UserScatteredDataParallel, train.py:184-186) joins the job: this process takes GPU LOCAL_RANK
    and the default process group is created (NCCL; gloo when there is no CUDA device, i.e. in the CPU tests).
    Returns (rank, world). A no-op without the launcher's variables or when a group already exists."""
    import os
    import torch.distributed as dist
    if not dist.is_available():
        return 0, 1
    if not dist.is_initialized():
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if world <= 1 or "RANK" not in os.environ:
            return 0, 1
        use_cuda = torch.cuda.is_available() and torch.cuda.device_count() > 0
        if use_cuda:
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", os.environ["RANK"])) % torch.cuda.device_count())
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend="nccl" if use_cuda else "gloo")
        _guard_checkpoint_writes()
    return dist.get_rank(), dist.get_world_size()


_save_guarded = False


def _guard_checkpoint_writes():
    """Every rank runs the same unmodified script, so every rank would `torch.save` the same
    cfg.DIR/{encoder,decoder,history}_epoch_N.pth at the same time (train.py:139-152) - concurrent non-atomic writes to one
    path. The replicas hold identical weights, so once this process has joined a job as rank r > 0 its torch.save() calls
    write nothing; rank 0 writes, and all ranks meet at a barrier after every save so nobody reads a half-written file.
    SSEG_CKPT_ALL_RANKS=1 restores per-rank writes (the tests use it with per-rank working directories)."""
    global _save_guarded
    import os
    import torch.distributed as dist
    if _save_guarded or os.environ.get("SSEG_CKPT_ALL_RANKS", "0") == "1" or dist.get_world_size() <= 1:
        return
    _save_guarded = True
    real_save = torch.save

    def save(obj, f, *args, **kwargs):
        try:
            if dist.get_rank() == 0:
                real_save(obj, f, *args, **kwargs)
        finally:
            if dist.is_initialized():
                dist.barrier()
    save.__wrapped__ = real_save
    torch.save = save


def _lift(out):
    """0-d tensors -> 1-d (reference dict_gather, data_parallel.py:33-36); recurse through dict/list/tuple."""
    if torch.is_tensor(out):
        return out.unsqueeze(0) if out.dim() == 0 else out
    if out is None:
        return None
    if isinstance(out, collections.abc.Mapping):
        return {k: _lift(v) for k, v in out.items()}
    if isinstance(out, collections.abc.Sequence):
        return type(out)(_lift(v) for v in out)
    return out


_copy_streams = {}


def _copy_stream(device):
    if device not in _copy_streams:
        _copy_streams[device] = cuda.Stream(device)
    return _copy_streams[device]


class DictGatherDataParallel(DataParallel):
    def gather(self, outputs, output_device):
        return _lift(outputs[0]) if len(outputs) == 1 else super().gather(outputs, output_device)


class UserScatteredDataParallel(DictGatherDataParallel):
    def __init__(self, module, device_ids=None, output_device=None, dim=0):
        # One process drives one GPU: keep only this process's device so nn.DataParallel never replicates.
        _, world = ensure_process_group()    # under torchrun: take GPU LOCAL_RANK and join the job (before the script's .cuda())
        if world == 1 and device_ids is not None and len(device_ids) > 1:
            # The reference drives all GPUs from ONE process (replicate + one Python thread per GPU, data_parallel.py:53-62).
            # This engine is one process per GPU; running on would silently train on 1/G of every batch list, so refuse.
            raise RuntimeError(
                "UserScatteredDataParallel(device_ids=%s) in a single process: the B200 engine runs one process per GPU. "
                "Launch the same script as  python -m torch.distributed.run --nnodes=1 --nproc-per-node %d "
                "--master-addr 127.0.0.1 <script> --gpus 0-%d ...  (the script itself needs no change: wrapping the model "
                "joins the job)." % (list(device_ids), len(device_ids), len(device_ids) - 1))
        if torch.cuda.is_available():
            dev = torch.cuda.current_device()
            super().__init__(module, device_ids=[dev], output_device=dev, dim=dim)
        else:
            nn.Module.__init__(self)
            self.module, self.device_ids, self.dim, self.output_device = module, [], dim, None
        self.requested_device_ids = device_ids

    def scatter(self, inputs, kwargs, device_ids):
        assert len(inputs) == 1 and len(kwargs) == 0
        batches = inputs[0]
        assert type(batches) in (tuple, list)
        rank, world = _rank_world()
        # the loader yields one dict per GPU of the job; a single-process run consumes entry `rank % len`
        mine = batches[rank % len(batches)]
        if isinstance(mine, dict) and 'skipped_for_rank' in mine:
            raise RuntimeError("the loader skipped this rank's entry: the per-GPU list has %d entries for a job of %d ranks - "
                               "launch with --gpus 0-%d (one entry per rank) or set SSEG_SHARD_LOADER=0"
                               % (len(batches), world, world - 1))
        dev = device_ids[0]
        with cuda.device(dev):
            main = cuda.current_stream()
            side = _copy_stream(dev)
            with cuda.stream(side):
                mine = async_copy_to(mine, dev, main_stream=main)
            main.wait_stream(side)
        return [[mine]], [{}]

    def forward(self, *inputs, **kwargs):
        if not self.device_ids:
            return _lift(self.module(*inputs, **kwargs))
        inputs, kwargs = self.scatter(inputs, kwargs, self.device_ids)
        return _lift(self.module(*inputs[0], **kwargs[0]))


class DataParallelWithCallback(DataParallel):
    """nn.DataParallel that marks SyncBN modules as parallel (reference replicate.py:50-67)."""

    def __init__(self, module, device_ids=None, output_device=None, dim=0):
        super().__init__(module, device_ids=device_ids, output_device=output_device, dim=dim)
        _mark_parallel(module)


def _mark_parallel(module):
    for m in module.modules():
        if hasattr(m, '__data_parallel_replicate__'):
            m.__data_parallel_replicate__(None, 0)


def patch_replication_callback(data_parallel):
    """Reference replicate.py:70-94: after this call the SyncBN layers of `data_parallel.module` use the
    synchronised (pooled-statistics) formula while training."""
    assert isinstance(data_parallel, DataParallel)
    _mark_parallel(data_parallel.module)
