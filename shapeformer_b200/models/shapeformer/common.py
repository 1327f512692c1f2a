"""Token glue used around the hot path (reference shapeformer/models/shapeformer/common.py)."""
import numpy as np
import torch


def filter_end_tokens(indices, end_tokens=(8192, 4096)):
    """(L, tuple_n) numpy -> rows without any end token (common.py:50-55)."""
    end = np.asarray(end_tokens)[None, ...]
    return indices[(indices != end).all(axis=1), :]


def batch_sparse2dense(sparse, empty_ind, dense_res, return_flattened=False, dim=3):
    """Packed (K,3) [batch, raveled pos, val] -> dense (B, res, res, res) (common.py:171-189).  Host-side convenience for
    reference-style callers; the batched device version is ImplicitDecoder.tokens_to_dense."""
    _, counts = torch.unique_consecutive(sparse[:, 0], return_counts=True)
    B = len(counts)
    dense = torch.full((B, dense_res ** dim), int(empty_ind), dtype=sparse.dtype, device=sparse.device)
    dense[sparse[:, 0], sparse[:, 1]] = sparse[:, 2]
    return dense if return_flattened else dense.view(B, *((dense_res,) * dim))
