"""geoutil.array2mesh of the reference (xgutils/geoutil.py:175-233) with the iso-surface extraction on the GPU
(csrc/mesh_kernels.cu).  The reference calls PyMCubes' marching_cubes; PyMCubes is not part of the reference tree, so the
extractor here is marching tetrahedra on the Kuhn subdivision: the same level set on the same grid, watertight, but a
different triangulation — vertex / face arrays are not comparable element-wise with PyMCubes' output."""
import numpy as np
import torch

from .. import _lib
from . import nputil


def iso_mesh(grid, thresh):
    """grid (R, R, R) fp32 CUDA tensor -> (verts (V, 3) fp32 in grid-index coordinates, faces (F, 3) int32), on the device."""
    lib = _lib.load()
    if not grid.is_cuda:
        raise _lib.Sfb200Error("iso_mesh needs a CUDA tensor (no CPU fallback)")
    g = grid.to(torch.float32).contiguous()
    R = g.shape[0]
    if g.dim() != 3 or g.shape[1] != R or g.shape[2] != R:
        raise _lib.Sfb200Error("grid must be (R, R, R)")
    dev = g.device
    with torch.cuda.device(dev):
        st = _lib.stream_ptr()
        flag = torch.empty(R ** 3 * 7, dtype=torch.int32, device=dev)
        _lib.check(lib.sfb200_mesh_mark_edges(_lib.ptr(g), R, float(thresh), _lib.ptr(flag), st), "sfb200_mesh_mark_edges")
        inc = torch.cumsum(flag, 0, dtype=torch.int32)
        vid = (inc - flag).contiguous()
        count = torch.empty((R - 1) ** 3, dtype=torch.int32, device=dev)
        _lib.check(lib.sfb200_mesh_count_faces(_lib.ptr(g), R, float(thresh), _lib.ptr(count), st), "sfb200_mesh_count_faces")
        finc = torch.cumsum(count, 0, dtype=torch.int32)
        foff = (finc - count).contiguous()
        V, F = int(inc[-1]), int(finc[-1])          # one small D2H sync: the mesh size is data dependent
        verts = torch.empty(V, 3, dtype=torch.float32, device=dev)
        faces = torch.empty(F, 3, dtype=torch.int32, device=dev)
        if V > 0:
            _lib.check(lib.sfb200_mesh_emit_vertices(_lib.ptr(g), R, float(thresh), _lib.ptr(flag), _lib.ptr(vid), _lib.ptr(verts), st),
                       "sfb200_mesh_emit_vertices")
        if F > 0:
            _lib.check(lib.sfb200_mesh_emit_faces(_lib.ptr(g), R, float(thresh), _lib.ptr(vid), _lib.ptr(foff), _lib.ptr(verts),
                                                  _lib.ptr(faces), st), "sfb200_mesh_emit_faces")
    return verts, faces


def array2mesh(array, thresh=0., dim=3, coords=None, bbox=np.array([[-1, -1, -1], [1, 1, 1]]), return_coords=False,
               if_decimate=False, decimate_face=4096, cart_coord=True, gaussian_sigma=None, device=None):
    """Same contract as the reference for dim = 3: 1-D array of R^3 values (numpy or torch; a CUDA tensor stays on the device)
    -> (verts (V, 3) float64 numpy scaled to the bounding box of `coords` / `bbox`, faces (F, 3) int numpy)."""
    if dim != 3:
        raise NotImplementedError("only dim = 3 (meshes) is on the B200 path")
    if if_decimate or gaussian_sigma is not None:
        raise NotImplementedError("decimation (igl) and smoothing (mcubes.smooth) are third-party post-processing, not provided")
    t = array if torch.is_tensor(array) else torch.from_numpy(np.asarray(array))
    if not t.is_cuda:
        t = t.to(device if device is not None else "cuda")
    R = int(np.ceil(np.power(t.numel(), 1.0 / 3)))
    verts, faces = iso_mesh(t.reshape(R, R, R), thresh)
    verts = verts.double().cpu().numpy()
    faces = faces.cpu().numpy().astype(int)
    if not cart_coord:
        verts = verts[:, [1, 0, 2]]
    verts = verts / (R - 1)
    if coords is not None:
        c = coords.detach().cpu().numpy() if torch.is_tensor(coords) else np.asarray(coords)
        bbmin, bbmax = c.reshape(-1, c.shape[-1]).min(0), c.reshape(-1, c.shape[-1]).max(0)
    else:
        bbmin, bbmax = np.asarray(bbox)[0], np.asarray(bbox)[1]
        coords = nputil.makeGrid(bb_min=bbmin, bb_max=bbmax, shape=(R, R, R))
    verts = verts * (bbmax - bbmin) + bbmin
    return (verts, faces, coords) if return_coords else (verts, faces)
