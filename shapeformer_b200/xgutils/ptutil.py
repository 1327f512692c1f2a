"""ravel / unravel (reference xgutils/ptutil.py:357-378) and tensor -> numpy conversion."""
import torch


def ravel_index(t, shape):
    if t.shape[-1] == 2:
        return t[..., 0] * shape[1] + t[..., 1]
    if t.shape[-1] == 3:
        return (t[..., 0] * shape[1] + t[..., 1]) * shape[2] + t[..., 2]
    raise ValueError("shape must be 2 or 3 dimensional")


def unravel_index(t, shape):
    if len(shape) == 2:
        return torch.stack([t // shape[1], t % shape[1]], -1)
    if len(shape) == 3:
        s12 = shape[1] * shape[2]
        return torch.stack([t // s12, t % s12 // shape[2], t % s12 % shape[2]], -1)
    raise ValueError("shape must be 2 or 3 dimensional")


def ths2nps(obj):
    if isinstance(obj, torch.Tensor):
        return obj.detach().cpu().numpy()
    if isinstance(obj, dict):
        return {k: ths2nps(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(ths2nps(v) for v in obj)
    return obj
