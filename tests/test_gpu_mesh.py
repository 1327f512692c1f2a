"""GPU iso-surface extraction (csrc/mesh_kernels.cu, shapeformer_b200.xgutils.geoutil.array2mesh) against its CPU restatement
(oracle/mesh_oracle.py) and against geometric properties of the level set — PyMCubes, which the reference calls, is absent
(parity unpinned, see the oracle's header)."""
import numpy as np
import pytest
import torch

from oracle import mesh_oracle
from shapeformer_b200.xgutils import geoutil

pytestmark = pytest.mark.gpu


def _sphere(R, r, c=(0.05, -0.02, 0.03)):
    ax = np.linspace(-1, 1, R)
    x, y, z = np.meshgrid(ax, ax, ax, indexing="ij")
    d = np.sqrt((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2)
    return (1.0 / (1.0 + np.exp((d - r) * 12.0))).astype(np.float32)      # occupancy-like: > 0.5 inside


@pytest.mark.parametrize("seed", [0, 1])
def test_iso_mesh_matches_cpu_restatement(cuda, seed):
    rng = np.random.default_rng(seed)
    R = 11
    g = _sphere(R, 0.55) + 0.25 * rng.standard_normal((R, R, R)).astype(np.float32)
    v, f = geoutil.iso_mesh(torch.from_numpy(g).to(cuda), 0.5)
    ov, of = mesh_oracle.iso_mesh(g, 0.5)
    assert v.shape == ov.shape and np.array_equal(v.cpu().numpy(), ov)          # bit-exact vertex positions, same order
    f = f.cpu().numpy()
    assert f.shape == of.shape and np.array_equal(np.sort(f, 1), np.sort(of, 1))
    same = (f == of).all(1).mean()
    assert same > 0.995, same                                                     # orientation (ties on degenerate triangles aside)


def test_sphere_is_watertight_oriented_and_encloses_the_right_volume(cuda):
    R, r = 64, 0.6
    g = _sphere(R, r)
    verts, faces = geoutil.array2mesh(torch.from_numpy(g).to(cuda).reshape(-1), thresh=0.5, dim=3)
    assert verts.dtype == np.float64 and faces.dtype.kind == "i" and verts.shape[1] == 3 and faces.shape[1] == 3
    e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]])
    und, cnt = np.unique(np.sort(e, 1), axis=0, return_counts=True)
    assert (cnt == 2).all()                                                      # closed 2-manifold
    assert len(np.unique(e, axis=0)) == len(e)                                   # every directed edge once: consistent orientation
    assert verts.shape[0] - len(und) + faces.shape[0] == 2                       # Euler characteristic of a sphere
    p = verts[faces]
    vol = np.einsum("ij,ij->i", p[:, 0], np.cross(p[:, 1], p[:, 2])).sum() / 6.0
    assert abs(vol - 4.0 / 3.0 * np.pi * r ** 3) < 0.02 * 4.0 / 3.0 * np.pi * r ** 3, vol      # outward normals: positive volume
    rad = np.linalg.norm(verts - np.array([0.05, -0.02, 0.03]), axis=1)
    assert np.abs(rad - r).max() < 2.0 / (R - 1)


def test_array2mesh_contract(cuda):
    g = _sphere(20, 0.5)
    coords = np.stack(np.meshgrid(*[np.linspace(-2, 2, 20)] * 3, indexing="ij"), -1).reshape(-1, 3)
    v, f, c = geoutil.array2mesh(g.reshape(-1), thresh=0.5, coords=coords, return_coords=True, device=cuda)
    assert c is coords and v.min() > -2 and v.max() < 2 and np.abs(np.linalg.norm(v - np.array([0.1, -0.04, 0.06]), axis=1) - 1.0).max() < 0.25
    v2, f2 = geoutil.array2mesh(np.zeros(8 ** 3, dtype=np.float32), thresh=0.5, device=cuda)
    assert v2.shape == (0, 3) and f2.shape == (0, 3)
    with pytest.raises(NotImplementedError):
        geoutil.array2mesh(g.reshape(-1), dim=2, device=cuda)
