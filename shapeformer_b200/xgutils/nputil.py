"""makeGrid (reference xgutils/nputil.py:618-654)."""
import numpy as np


def makeGrid(bb_min=(0, 0, 0), bb_max=(1, 1, 1), shape=(10, 10, 10), mode="on", flatten=True, indexing="ij"):
    bb_min, bb_max = np.asarray(bb_min, dtype=np.float64), np.asarray(bb_max, dtype=np.float64)
    if isinstance(shape, int):
        shape = [shape] * len(bb_min)
    axes = []
    for i, n in enumerate(shape):
        if mode == "on":
            axes.append(np.linspace(bb_min[i], bb_max[i], n))
        elif mode == "in":
            off = (bb_max[i] - bb_min[i]) / 2.0 / n
            axes.append(np.linspace(bb_min[i] + off, bb_max[i] - off, n))
        else:
            raise ValueError(mode)
    grid = np.stack(np.meshgrid(*axes, sparse=False, indexing=indexing), axis=-1)
    return grid.reshape(-1, grid.shape[-1]) if flatten else grid


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))
