"""CPU: the marching-tetrahedra restatement (oracle/mesh_oracle.py) is pinned against geometric properties of the level set —
PyMCubes, which the reference calls (xgutils/geoutil.py:208), is absent, so there is nothing element-wise to compare with."""
import numpy as np

from oracle import mesh_oracle


def test_oracle_sphere_is_closed_oriented_and_has_the_right_volume():
    R, r = 16, 0.6
    ax = np.linspace(-1, 1, R)
    x, y, z = np.meshgrid(ax, ax, ax, indexing="ij")
    g = (1.0 / (1.0 + np.exp((np.sqrt(x * x + y * y + z * z) - r) * 12.0))).astype(np.float32)
    v, f = mesh_oracle.iso_mesh(g, 0.5)
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
    und, cnt = np.unique(np.sort(e, 1), axis=0, return_counts=True)
    assert (cnt == 2).all() and len(np.unique(e, axis=0)) == len(e)
    assert v.shape[0] - len(und) + f.shape[0] == 2
    p = (v / (R - 1) * 2 - 1)[f].astype(np.float64)
    vol = np.einsum("ij,ij->i", p[:, 0], np.cross(p[:, 1], p[:, 2])).sum() / 6.0
    assert abs(vol - 4 / 3 * np.pi * r ** 3) < 0.05 * 4 / 3 * np.pi * r ** 3


def test_oracle_empty_and_full_grids_give_no_mesh():
    for val in (0.0, 1.0):
        v, f = mesh_oracle.iso_mesh(np.full((5, 5, 5), val, dtype=np.float32), 0.5)
        assert v.shape == (0, 3) and f.shape == (0, 3)
