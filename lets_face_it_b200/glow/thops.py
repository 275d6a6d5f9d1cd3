"""Tensor helpers of the reference's `glow/thops.py` (sum / mean over several dims, feature split / concat).
Pure view/indexing glue on the host side; the fused kernels do the same splits in registers."""
import torch


def _reduce(fn, tensor, dim, keepdim):
    if dim is None:
        return fn(tensor)
    dims = sorted([dim] if isinstance(dim, int) else list(dim))
    return fn(tensor, dim=dims, keepdim=keepdim)


def sum(tensor, dim=None, keepdim=False):
    return _reduce(torch.sum, tensor, dim, keepdim)


def mean(tensor, dim=None, keepdim=False):
    return _reduce(torch.mean, tensor, dim, keepdim)


def split_feature(tensor, type="split"):
    """"split": first half / second half of the channel dim; "cross": even / odd channels (thops.py:36-44)."""
    C = tensor.size(1)
    if type == "split":
        return tensor[:, : C // 2, ...], tensor[:, C // 2 :, ...]
    if type == "cross":
        return tensor[:, 0::2, ...], tensor[:, 1::2, ...]
    raise ValueError(type)


def cat_feature(tensor_a, tensor_b):
    return torch.cat((tensor_a, tensor_b), dim=1)
