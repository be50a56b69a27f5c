"""Host -> device input pipeline with the copy hidden behind the previous step.

The reference moves every batch with a blocking `async_copy_to` right before the forward pass (train.py:48-50,
lib/nn/parallel/data_parallel.py:82-96) after its loader workers produced normalised fp32 tensors
(mit_semseg/dataset.py:53-63). Here one batch is always in flight:

    loader (raw=True: uint8 HWC images + uint8 label maps, 4x fewer bytes)
      -> pinned staging buffers (reused per batch shape)
      -> cudaMemcpyAsync on a COPY stream
      -> sseg_image_transform / sseg_label_transform on the same stream  (the reference's img_transform / segm_transform,
         bit-identical, csrc/input.cu)
      -> event; `next()` makes the compute stream wait on it and hands out {'img_data', 'seg_label'} on the device.

Two slots alternate, so batch i+1 is copied and transformed while step i still reads batch i. Batches of the reference's own
format (float `img_data`, int64 `seg_label`) are accepted too: they skip the transform kernels and are only staged and copied.
The host side of "one batch in flight" - pulling the next batch out of the loader, the copy into pinned staging (skipped
for tensors that are pinned already) and enqueueing the transfer - runs on a helper thread started when a batch is handed
out, i.e. WHILE the caller launches its training step; measured on B200 without it (profiles/r2_first_run.log) the staging
sat between two steps and the prefetcher was slower than the plain synchronous copy.
Real ADE20K batches change shape from step to step; buffers are cached per shape, the step programs (engine/functional.py)
are cached per shape as well.
"""
import torch

from . import ops


class _Slot:
    def __init__(self):
        self.bufs = {}      # (name, shape, dtype) -> (pinned host tensor, device tensor)
        self.keep_alive = {}
        self.event = None
        self.feed = None


class DevicePrefetcher:
    _use_streams = True     # tests on the emulated / simulated ABI run the same code without CUDA streams

    def __init__(self, loader, device=None, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225), background=True):
        self.loader = loader
        self.dev = torch.device(device if device is not None else ("cuda:%d" % torch.cuda.current_device()))
        self.mean, self.std = tuple(mean), tuple(std)
        self.slots = [_Slot(), _Slot()]
        self.turn = 0
        self.h2d_bytes = 0          # bytes copied for the batch handed out last (bench bookkeeping)
        self.copy_stream = torch.cuda.Stream(self.dev) if self._use_streams else None
        self.background = background
        self._it = None
        self._pending = None
        self._thread = None

    # ------------------------------------------------------------------------------------------ staging
    def _staged(self, slot, name, host):
        """host tensor -> device tensor through a pinned buffer of the same shape (both cached in the slot)"""
        key = (name, tuple(host.shape), host.dtype)
        if key not in slot.bufs:
            pinned = torch.empty(host.shape, dtype=host.dtype)
            if self._use_streams and not host.is_pinned():
                pinned = pinned.pin_memory()
            slot.bufs[key] = (pinned, torch.empty(host.shape, dtype=host.dtype, device=self.dev))
        pinned, dev = slot.bufs[key]
        if host.is_pinned():
            dev.copy_(host, non_blocking=True)      # the loader (pin_memory=True) already produced page-locked memory
            slot.keep_alive[key] = host             # ... which must outlive the asynchronous copy
        else:
            pinned.copy_(host)
            dev.copy_(pinned, non_blocking=True)
        self._bytes += host.numel() * host.element_size()
        return dev

    def _device_out(self, slot, name, shape, dtype):
        key = (name, tuple(shape), dtype)
        if key not in slot.bufs:
            slot.bufs[key] = (None, torch.empty(shape, dtype=dtype, device=self.dev))
        return slot.bufs[key][1]

    def _issue(self, batch):
        """start copying + transforming `batch` into the next slot (on the copy stream)"""
        if isinstance(batch, (list, tuple)):     # the reference's collate hands one dict per GPU: this process owns one GPU
            assert len(batch) == 1, "one process per GPU: the loader must yield this rank's batch only"
            batch = batch[0]
        slot = self.slots[self.turn]
        self.turn ^= 1
        self._bytes = 0
        if slot.event is not None:
            slot.event.synchronize()     # the pinned buffers of this slot are rewritten by the host below (long done)

        def work():
            if 'img_u8' in batch:
                u8 = self._staged(slot, 'img_u8', batch['img_u8'])
                seg = self._staged(slot, 'seg_u8', batch['seg_u8'])
                valid = self._staged(slot, 'valid_hw', batch['valid_hw'])
                n, h, w, _ = u8.shape
                img = self._device_out(slot, 'img_data', (n, 3, h, w), torch.float32)
                lab = self._device_out(slot, 'seg_label', tuple(seg.shape), torch.int64)
                ops.image_transform(u8, valid, img, self.mean, self.std)
                ops.label_transform(seg, valid, int(batch['segm_downsampling_rate']), lab)
                return {'img_data': img, 'seg_label': lab}
            return {k: (self._staged(slot, k, v) if torch.is_tensor(v) else v) for k, v in batch.items()}

        if self._use_streams:
            with torch.cuda.stream(self.copy_stream):
                slot.feed = work()
                slot.event = torch.cuda.Event()
                slot.event.record(self.copy_stream)
        else:
            slot.feed = work()
        slot.nbytes = self._bytes
        return slot

    # ------------------------------------------------------------------------------------------ iteration
    def _fetch_and_issue(self):
        """next batch of the loader -> staged, copy / transform enqueued; None at the end of the loader"""
        try:
            batch = next(self._it)
        except StopIteration:
            return None
        return self._issue(batch)

    def _start_next(self):
        if not self._use_streams or not self.background:
            self._pending, self._thread = self._fetch_and_issue(), None
            return
        import threading
        box = {}
        dev_index = self.dev.index if self.dev.index is not None else torch.cuda.current_device()

        def work():
            try:
                torch.cuda.set_device(dev_index)
                box["slot"] = self._fetch_and_issue()
            except BaseException as exc:   # noqa: BLE001 - re-raised in the consumer's thread
                box["exc"] = exc
        self._thread = threading.Thread(target=work, daemon=True)
        self._box = box
        self._thread.start()

    def _collect(self):
        if self._thread is not None:
            self._thread.join()
            self._thread = None
            if "exc" in self._box:
                raise self._box["exc"]
            self._pending = self._box.get("slot")
        return self._pending

    def __iter__(self):
        self._it = iter(self.loader)
        self._pending, self._thread = None, None
        self._start_next()
        return self

    def __next__(self):
        if self._it is None:
            iter(self)
        slot = self._collect()
        if slot is None:
            raise StopIteration
        if self._use_streams:
            torch.cuda.current_stream(self.dev).wait_event(slot.event)
        self.h2d_bytes = slot.nbytes
        feed = slot.feed
        # the slot handed out two calls ago is about to be overwritten: its consumer (the step launched after that call) was
        # enqueued on the compute stream before this point, so order the copy stream behind it
        if self._use_streams:
            self.copy_stream.wait_stream(torch.cuda.current_stream(self.dev))
        self._start_next()     # runs while the caller launches its step on the batch returned here
        return feed
