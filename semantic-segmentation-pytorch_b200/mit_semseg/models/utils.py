"""Weight-file helper with the reference's name (mit_semseg/models/utils.py:10-18 downloads ImageNet weights).

This environment has no network, so `load_url` only resolves files that are already present under `model_dir`
and raises a clear error otherwise."""
import os

import torch


def load_url(url, model_dir='./pretrained', map_location=None):
    filename = url.split('/')[-1]
    cached_file = os.path.join(model_dir, filename)
    if not os.path.exists(cached_file):
        raise FileNotFoundError(
            "pretrained weights %s are not cached at %s and this build never downloads; pass weights=<file> to "
            "ModelBuilder.build_encoder or place the file there" % (url, cached_file))
    return torch.load(cached_file, map_location=map_location)
