"""`as_numpy` & co. (mit_semseg/lib/utils/th.py:8-41): walk nested containers of tensors. The reference's versions predate
torch 0.4 (`Variable`, `volatile`); on a current torch they reduce to what is written here."""
import collections.abc

import numpy as np
import torch

__all__ = ['as_variable', 'as_numpy', 'mark_volatile']


def _walk(obj, leaf):
    if torch.is_tensor(obj):
        return leaf(obj)
    if isinstance(obj, collections.abc.Mapping):
        return {k: _walk(v, leaf) for k, v in obj.items()}
    if isinstance(obj, collections.abc.Sequence) and not isinstance(obj, (str, bytes)):
        return [_walk(v, leaf) for v in obj]
    return obj


def as_variable(obj):
    """tensors are their own Variables since torch 0.4"""
    return _walk(obj, lambda t: t)


def as_numpy(obj):
    """tensors -> numpy arrays (on the host), containers rebuilt as dict / list, everything else -> np.array(obj)"""
    if torch.is_tensor(obj):
        return obj.detach().cpu().numpy()
    if isinstance(obj, collections.abc.Mapping):
        return {k: as_numpy(v) for k, v in obj.items()}
    if isinstance(obj, collections.abc.Sequence) and not isinstance(obj, (str, bytes)):
        return [as_numpy(v) for v in obj]
    return np.array(obj)


def mark_volatile(obj):
    """`volatile` is gone: inference code runs under torch.no_grad(); tensors are returned detached"""
    return _walk(obj, lambda t: t.detach())
