"""ADE20K-style datasets behind the reference's interface (mit_semseg/dataset.py:22-296): `TrainDataset`, `ValDataset`,
`TestDataset` take the same constructor arguments (`.odgt` list or list of records, the `cfg.DATASET` node) and return the
same dictionaries - bit for bit under the same numpy seed, because the draws happen in the reference's order:
seed(index) + shuffle on the first item, one short-edge choice per batch, one flip coin per image, a reshuffle whenever the
cursor wraps (tests/test_dataset.py compares against the imported reference on synthetic image files).

B200 addition: `raw=True` keeps the decode / resize / flip on the host (PIL) but returns UNNORMALISED bytes -

    img_u8     uint8 [B, H, W, 3]   (HWC as decoded, zero padded to the batch size)
    seg_u8     uint8 [B, H/r, W/r]  (class ids 0..150 as stored, zero padded)
    valid_hw   int32 [B, 2]         (rows, columns of every image that are real; labels: valid // r rounded up)

- a quarter of the fp32 bytes over PCIe; `engine.prefetch.DevicePrefetcher` then runs the reference's `img_transform` /
`segm_transform` ((x/255 - mean)/std to fp32 NCHW, label - 1 to int64, the padding written as zeros exactly like the
reference's pre-zeroed batch tensors) as one kernel each on the copy stream (csrc/elementwise.cu::image_transform_kernel).
"""
import json
import os

import numpy as np
import torch
from PIL import Image

IMAGE_MEAN = (0.485, 0.456, 0.406)
IMAGE_STD = (0.229, 0.224, 0.225)

_RESAMPLE = {'nearest': Image.NEAREST, 'bilinear': Image.BILINEAR, 'bicubic': Image.BICUBIC}


def imresize(im, size, interp='bilinear'):
    if interp not in _RESAMPLE:
        raise Exception('resample method undefined!')
    return im.resize(size, _RESAMPLE[interp])


def _ceil_to(x, p):
    """smallest multiple of p that is >= x"""
    return ((x - 1) // p + 1) * p


def _fit_scale(h, w, short, longest):
    """resize factor that brings the short edge to `short` unless the long edge would pass `longest`"""
    return min(short / float(min(h, w)), longest / float(max(h, w)))


class BaseDataset(torch.utils.data.Dataset):
    def __init__(self, odgt, opt, raw=False, **kwargs):
        self.imgSizes = opt.imgSizes
        self.imgMaxSize = opt.imgMaxSize
        self.padding_constant = opt.padding_constant   # largest down-sampling rate of the network
        self.raw = bool(raw)
        self.parse_input_list(odgt, **kwargs)
        self._mean = torch.tensor(IMAGE_MEAN, dtype=torch.float32).view(3, 1, 1)
        self._std = torch.tensor(IMAGE_STD, dtype=torch.float32).view(3, 1, 1)

    def parse_input_list(self, odgt, max_sample=-1, start_idx=-1, end_idx=-1):
        if isinstance(odgt, list):
            records = odgt
        elif isinstance(odgt, str):
            with open(odgt, 'r') as f:
                records = [json.loads(line.rstrip()) for line in f]
        else:
            raise TypeError('odgt: a file name or a list of records')
        if max_sample > 0:
            records = records[0:max_sample]
        if start_idx >= 0 and end_idx >= 0:
            records = records[start_idx:end_idx]
        self.list_sample = records
        self.num_sample = len(records)
        assert self.num_sample > 0
        print('# samples: {}'.format(self.num_sample))

    def normalize(self, chw):
        return (chw - self._mean) / self._std

    def img_transform(self, img):
        """PIL RGB / HWC uint8 -> normalised float CHW (dataset.py:53-58)"""
        chw = (np.float32(np.array(img)) / 255.).transpose((2, 0, 1))
        return self.normalize(torch.from_numpy(chw.copy()))

    def segm_transform(self, segm):
        """stored ids 0..150 -> labels -1..149 (dataset.py:60-63)"""
        return torch.from_numpy(np.array(segm)).long() - 1

    def round2nearest_multiple(self, x, p):
        return _ceil_to(x, p)

    def _scaled_inputs(self, img):
        """one resized, transformed copy of `img` per entry of imgSizes, each [1, 3, h, w] with h, w multiples of the
        padding constant (evaluation: dataset.py:214-232, 265-283)"""
        ow, oh = img.size
        out = []
        for short in self.imgSizes:
            s = _fit_scale(oh, ow, short, self.imgMaxSize)
            th, tw = _ceil_to(int(oh * s), self.padding_constant), _ceil_to(int(ow * s), self.padding_constant)
            out.append(self.img_transform(imresize(img, (tw, th), interp='bilinear')).unsqueeze(0).contiguous())
        return out


def _job():
    """(rank, world) of the one-process-per-GPU job this loader feeds, from the launcher's environment (it is also what the
    DataLoader's worker processes see); (0, 1) outside such a job or with SSEG_SHARD_LOADER=0."""
    if os.environ.get("SSEG_SHARD_LOADER", "1") == "0":
        return 0, 1
    try:
        world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    except ValueError:
        return 0, 1
    return (rank, world) if world > 1 else (0, 1)


def _other_ranks_entry(index):
    """train.py builds DataLoader(batch_size=len(gpus), shuffle=False): item `index` is the batch of GPU index % len(gpus),
    and the documented launch has len(gpus) == WORLD_SIZE (`--gpus 0-(N-1)` under torch.distributed.run)."""
    rank, world = _job()
    return world > 1 and index % world != rank


class TrainDataset(BaseDataset):
    """One item = one per-GPU batch of `batch_per_gpu` images of the same orientation, resized to a common random short
    edge and padded to a common size (dataset.py:70-186). `len()` is the reference's fake 1e10: every loader worker walks
    its own shuffled copy of the list."""

    def __init__(self, root_dataset, odgt, opt, batch_per_gpu=1, **kwargs):
        super().__init__(odgt, opt, **kwargs)
        self.root_dataset = root_dataset
        self.segm_downsampling_rate = opt.segm_downsampling_rate
        self.batch_per_gpu = batch_per_gpu
        self.batch_record_list = [[], []]    # pending records: [portrait (h > w), landscape / square]
        self.cur_idx = 0
        self.if_shuffled = False

    def _get_sub_batch(self):
        pending = self.batch_record_list
        while True:
            rec = self.list_sample[self.cur_idx]
            bucket = 0 if rec['height'] > rec['width'] else 1
            pending[bucket].append(rec)
            self.cur_idx += 1
            if self.cur_idx >= self.num_sample:
                self.cur_idx = 0
                np.random.shuffle(self.list_sample)
            for b in (0, 1):    # at most one bucket can have filled up: only one record was added
                if len(pending[b]) == self.batch_per_gpu:
                    full, pending[b] = pending[b], []
                    return full

    def __getitem__(self, index):
        if not self.if_shuffled:     # the first item decides this worker's order (a shuffle in __init__ would be shared)
            np.random.seed(index)
            np.random.shuffle(self.list_sample)
            self.if_shuffled = True
        records = self._get_sub_batch()
        B, rate, pad = self.batch_per_gpu, self.segm_downsampling_rate, self.padding_constant
        assert pad >= rate, 'padding constant must be equal or large than segm downsamping rate'
        short = np.random.choice(self.imgSizes) if isinstance(self.imgSizes, (list, tuple)) else self.imgSizes

        widths, heights = np.zeros(B, np.int32), np.zeros(B, np.int32)
        for i, rec in enumerate(records):
            s = min(short / min(rec['height'], rec['width']), self.imgMaxSize / max(rec['height'], rec['width']))
            widths[i], heights[i] = rec['width'] * s, rec['height'] * s          # truncated by the int32 store
        bw, bh = int(_ceil_to(np.max(widths), pad)), int(_ceil_to(np.max(heights), pad))

        if _other_ranks_entry(index):
            # One process per GPU: every rank runs the same loader (same seeds, same order) and keeps entry `rank` of each
            # per-GPU list (lib/nn/parallel.py::scatter). Entries that belong to other ranks are not decoded / resized /
            # normalised here - only their random draws are consumed, so the streams of all ranks stay identical.
            for _ in records:
                np.random.choice([0, 1])
            return {'skipped_for_rank': index % _job()[1]}

        if self.raw:
            images = torch.zeros(B, bh, bw, 3, dtype=torch.uint8)
            segms = torch.zeros(B, bh // rate, bw // rate, dtype=torch.uint8)
            valid = torch.zeros(B, 2, dtype=torch.int32)
        else:
            images = torch.zeros(B, 3, bh, bw)
            segms = torch.zeros(B, bh // rate, bw // rate).long()

        for i, rec in enumerate(records):
            img = Image.open(os.path.join(self.root_dataset, rec['fpath_img'])).convert('RGB')
            segm = Image.open(os.path.join(self.root_dataset, rec['fpath_segm']))
            assert segm.mode == "L"
            assert img.size[0] == segm.size[0] and img.size[1] == segm.size[1]
            if np.random.choice([0, 1]):
                img, segm = img.transpose(Image.FLIP_LEFT_RIGHT), segm.transpose(Image.FLIP_LEFT_RIGHT)
            size = (widths[i], heights[i])     # every sample of the batch has its own scale
            img, segm = imresize(img, size, interp='bilinear'), imresize(segm, size, interp='nearest')
            # label map: onto a canvas whose size divides by the rate (so the strided pick lines up), then every rate-th
            canvas = Image.new('L', (_ceil_to(segm.size[0], rate), _ceil_to(segm.size[1], rate)), 0)
            canvas.paste(segm, (0, 0))
            segm = imresize(canvas, (canvas.size[0] // rate, canvas.size[1] // rate), interp='nearest')
            if self.raw:
                a, s = np.array(img), np.array(segm)
                images[i, :a.shape[0], :a.shape[1]] = torch.from_numpy(a.copy())
                segms[i, :s.shape[0], :s.shape[1]] = torch.from_numpy(s.copy())
                valid[i, 0], valid[i, 1] = a.shape[0], a.shape[1]
            else:
                t, s = self.img_transform(img), self.segm_transform(segm)
                images[i][:, :t.shape[1], :t.shape[2]] = t
                segms[i][:s.shape[0], :s.shape[1]] = s

        if self.raw:
            return {'img_u8': images, 'seg_u8': segms, 'valid_hw': valid, 'segm_downsampling_rate': rate}
        return {'img_data': images, 'seg_label': segms}

    def __len__(self):
        return int(1e10)


class ValDataset(BaseDataset):
    def __init__(self, root_dataset, odgt, opt, **kwargs):
        super().__init__(odgt, opt, **kwargs)
        self.root_dataset = root_dataset

    def __getitem__(self, index):
        rec = self.list_sample[index]
        img = Image.open(os.path.join(self.root_dataset, rec['fpath_img'])).convert('RGB')
        segm = Image.open(os.path.join(self.root_dataset, rec['fpath_segm']))
        assert segm.mode == "L"
        assert img.size[0] == segm.size[0] and img.size[1] == segm.size[1]
        return {'img_ori': np.array(img), 'img_data': self._scaled_inputs(img),
                'seg_label': self.segm_transform(segm).unsqueeze(0).contiguous(), 'info': rec['fpath_img']}

    def __len__(self):
        return self.num_sample


class TestDataset(BaseDataset):
    def __init__(self, odgt, opt, **kwargs):
        super().__init__(odgt, opt, **kwargs)

    def __getitem__(self, index):
        rec = self.list_sample[index]
        img = Image.open(rec['fpath_img']).convert('RGB')
        return {'img_ori': np.array(img), 'img_data': self._scaled_inputs(img), 'info': rec['fpath_img']}

    def __len__(self):
        return self.num_sample
