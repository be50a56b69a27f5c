"""Host-side helpers the reference's scripts import from `mit_semseg.utils` (train.py:15, eval.py / test.py): logging,
running averages, pixel accuracy, per-class intersection / union, colour rendering, `--gpus` parsing. Same names, arguments
and results as mit_semseg/utils.py:10-200 (tests/test_scripts_support.py compares with the imported reference); none of it
is on the timed path."""
import fnmatch
import logging
import os
import re
import sys

import numpy as np


def setup_logger(distributed_rank=0, filename="log.txt"):
    """stdout logger named "Logger"; ranks > 0 get the logger without a handler (utils.py:10-22)"""
    logger = logging.getLogger("Logger")
    logger.setLevel(logging.DEBUG)
    if distributed_rank > 0:
        return logger
    handler = logging.StreamHandler(stream=sys.stdout)
    handler.setLevel(logging.DEBUG)
    handler.setFormatter(logging.Formatter(
        "[%(asctime)s %(levelname)s %(filename)s line %(lineno)d %(process)d] %(message)s"))
    logger.addHandler(handler)
    return logger


def find_recursive(root_dir, ext='.jpg'):
    return [os.path.join(d, f) for d, _, names in os.walk(root_dir) for f in fnmatch.filter(names, '*' + ext)]


class AverageMeter(object):
    """weighted running mean; `val` is the last value"""

    def __init__(self):
        self.initialized = False
        self.val = self.avg = self.sum = self.count = None

    def initialize(self, val, weight):
        self.val, self.avg, self.sum, self.count, self.initialized = val, val, val * weight, weight, True

    def add(self, val, weight):
        self.val = val
        self.sum += val * weight
        self.count += weight
        self.avg = self.sum / self.count

    def update(self, val, weight=1):
        (self.add if self.initialized else self.initialize)(val, weight)

    def value(self):
        return self.val

    def average(self):
        return self.avg


def unique(ar, return_index=False, return_inverse=False, return_counts=False):
    """np.unique over the flattened input (the reference carries its own copy, utils.py:66-110; the results are numpy's)"""
    ar = np.asanyarray(ar).flatten()
    if not (return_index or return_inverse or return_counts):
        return np.unique(ar)
    return np.unique(ar, return_index=return_index, return_inverse=return_inverse, return_counts=return_counts)


def colorEncode(labelmap, colors, mode='RGB'):
    """label map [H, W] -> uint8 colour image [H, W, 3]; negative labels stay black (utils.py:113-128)"""
    labelmap = labelmap.astype('int')
    rgb = np.zeros(labelmap.shape[:2] + (3,), dtype=np.uint8)
    for label in np.unique(labelmap):
        if label >= 0:
            rgb[labelmap == label] = colors[label]
    return rgb[:, :, ::-1] if mode == 'BGR' else rgb


def accuracy(preds, label):
    """fraction of labelled pixels (label >= 0) predicted correctly, and their number (utils.py:131-136)"""
    valid = (label >= 0)
    valid_sum = valid.sum()
    return float((valid * (preds == label)).sum()) / (valid_sum + 1e-10), valid_sum


def intersectionAndUnion(imPred, imLab, numClass):
    """per-class intersection and union pixel counts; unlabeled ground truth (-1) is excluded from both (utils.py:139-160)"""
    pred = np.asarray(imPred).copy() + 1
    lab = np.asarray(imLab).copy() + 1
    pred = pred * (lab > 0)
    bins = dict(bins=numClass, range=(1, numClass))
    inter = np.histogram(pred * (pred == lab), **bins)[0]
    union = np.histogram(pred, **bins)[0] + np.histogram(lab, **bins)[0] - inter
    return inter, union


class NotSupportedCliException(Exception):
    pass


def process_range(xpu, inp):
    lo, hi = sorted(map(int, inp))
    return ('{}{}'.format(xpu, i) for i in range(lo, hi + 1))


_ONE = re.compile(r'^(?:gpu)?(\d+)$')
_SPAN = re.compile(r'^(?:gpu(\d+)-(?:gpu)?(\d+)|(\d+)-(\d+))$')


def parse_devices(input_devices):
    """'0-3' / '0,2' / 'gpu1-gpu2' -> ['gpu0', 'gpu1', ...] without duplicates (utils.py:185-200)"""
    out = []
    for item in input_devices.split(','):
        text = item.lower().strip()
        one, span = _ONE.match(text), _SPAN.match(text)
        if one:
            names = ['gpu%s' % one.group(1)]
        elif span:
            names = process_range('gpu', [g for g in span.groups() if g is not None])
        else:
            raise NotSupportedCliException('Can not recognize device: "{}"'.format(item))
        for n in names:
            if n not in out:
                out.append(n)
    return out
