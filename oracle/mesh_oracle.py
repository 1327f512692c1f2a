"""TEST INFRASTRUCTURE ONLY (imported by tests/ — never by shapeformer_b200/).

CPU restatement (numpy, plain loops — small grids only) of the iso-surface extractor of csrc/mesh_kernels.cu: marching tetrahedra
on the Kuhn subdivision, the stand-in for the step geoutil.array2mesh performs in the reference (xgutils/geoutil.py:175-233)
through PyMCubes.  PARITY UNPINNED: PyMCubes (pinned nowhere in the reference's environment.yml beyond `PyMCubes`) is a
third-party dependency absent from /root/reference and from this image, and its marching-cubes case tables cannot be restated
from the reference; the tests therefore pin this restatement against geometric properties of the level set (watertightness,
Euler characteristic, enclosed volume) instead of PyMCubes output."""
import numpy as np

DIRS = [(1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (1, 0, 1), (0, 1, 1), (1, 1, 1)]
TETS = [(0, 1, 3, 7), (0, 1, 5, 7), (0, 2, 3, 7), (0, 2, 6, 7), (0, 4, 5, 7), (0, 4, 6, 7)]
DIR_OF = {1: 0, 2: 1, 4: 2, 3: 3, 5: 4, 6: 5, 7: 6}


def _corner(code):
    return code & 1, (code >> 1) & 1, (code >> 2) & 1


def iso_mesh(grid, thresh):
    g = np.asarray(grid, dtype=np.float32)
    t = np.float32(thresh)
    R = g.shape[0]
    vid, verts = {}, []
    for i in range(R):
        for j in range(R):
            for k in range(R):
                for d, (a, b, c) in enumerate(DIRS):
                    i2, j2, k2 = i + a, j + b, k + c
                    if i2 < R and j2 < R and k2 < R and (g[i, j, k] > t) != (g[i2, j2, k2] > t):
                        va, vb = g[i, j, k], g[i2, j2, k2]
                        s = np.float32(t - va) / np.float32(vb - va)
                        vid[(i, j, k, d)] = len(verts)
                        verts.append((np.float32(i) + s * np.float32(a), np.float32(j) + s * np.float32(b), np.float32(k) + s * np.float32(c)))
    verts = np.array(verts, dtype=np.float32).reshape(-1, 3)
    faces = []
    for i in range(R - 1):
        for j in range(R - 1):
            for k in range(R - 1):
                ins = [bool(g[i + _corner(q)[0], j + _corner(q)[1], k + _corner(q)[2]] > t) for q in range(8)]
                for tet in TETS:
                    m = [ins[c] for c in tet]
                    cnt = sum(m)
                    if cnt in (0, 4):
                        continue

                    def ev(a, b):
                        a, b = min(a, b), max(a, b)
                        ca, cb = tet[a], tet[b]
                        x, y, z = _corner(ca)
                        return vid[(i + x, j + y, k + z, DIR_OF[cb - ca])]

                    ci = np.mean([_corner(tet[u]) for u in range(4) if m[u]], axis=0)
                    co = np.mean([_corner(tet[u]) for u in range(4) if not m[u]], axis=0)
                    if cnt in (1, 3):
                        lone = [u for u in range(4) if m[u] == (cnt == 1)][0]
                        tris = [[ev(u, lone) for u in range(4) if u != lone]]
                    else:
                        a = [u for u in range(4) if m[u]]
                        b = [u for u in range(4) if not m[u]]
                        v00, v01, v11, v10 = ev(a[0], b[0]), ev(a[0], b[1]), ev(a[1], b[1]), ev(a[1], b[0])
                        tris = [[v00, v01, v11], [v00, v11, v10]]
                    for tr in tris:
                        p = verts[tr].astype(np.float64)
                        n = np.cross(p[1] - p[0], p[2] - p[0])
                        if np.dot(n, co - ci) < 0:
                            tr = [tr[0], tr[2], tr[1]]
                        faces.append(tr)
    return verts, np.array(faces, dtype=np.int32).reshape(-1, 3)
