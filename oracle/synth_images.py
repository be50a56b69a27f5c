"""Synthetic ADE20K-style image / label files for the dataset parity tests (TEST INFRASTRUCTURE, see oracle/README or
DESIGN.md section 5). Lossless PNGs from a seeded generator, so the files - and therefore every resize / flip / crop the
datasets derive from them - are identical wherever they are generated."""
import os
from types import SimpleNamespace

import numpy as np
from PIL import Image

SIZES = [(97, 130), (120, 90), (64, 64), (150, 101), (88, 140), (141, 77), (100, 100), (75, 133)]   # (height, width)


def write_dataset(root, seed=1234):
    """-> list of odgt records (fpath_img, fpath_segm, width, height) for files written under `root`."""
    rng = np.random.RandomState(seed)
    os.makedirs(os.path.join(root, "images"), exist_ok=True)
    os.makedirs(os.path.join(root, "annotations"), exist_ok=True)
    records = []
    for i, (h, w) in enumerate(SIZES):
        # smooth-ish content (so bilinear resizing is not degenerate) + noise
        yy, xx = np.mgrid[0:h, 0:w]
        base = np.stack([127 + 100 * np.sin(xx / (5.0 + c) + i) * np.cos(yy / (7.0 + c)) for c in range(3)], axis=2)
        img = np.clip(base + rng.randint(-20, 21, size=(h, w, 3)), 0, 255).astype(np.uint8)
        segm = ((yy // 9 + xx // 11 + i) % 151).astype(np.uint8)          # ids 0..150, 0 = unlabeled
        fi, fs = "images/im_%02d.png" % i, "annotations/im_%02d.png" % i
        Image.fromarray(img, "RGB").save(os.path.join(root, fi))
        Image.fromarray(segm, "L").save(os.path.join(root, fs))
        records.append({"fpath_img": fi, "fpath_segm": fs, "width": w, "height": h})
    return records


def dataset_options(**over):
    opt = dict(imgSizes=(64, 80, 96), imgMaxSize=160, padding_constant=8, segm_downsampling_rate=8)
    opt.update(over)
    return SimpleNamespace(**opt)


def summarize(t):
    """compact fingerprint of a tensor: shape, float64 sums, a strided sample"""
    a = np.asarray(t.detach().cpu().numpy() if hasattr(t, "detach") else t)
    flat = a.astype(np.float64).ravel()
    return {"shape": np.array(a.shape, np.int64), "sum": np.array(flat.sum()), "abs": np.array(np.abs(flat).sum()),
            "sample": flat[::37][:4096].copy()}
