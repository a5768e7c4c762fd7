"""TEST INFRASTRUCTURE ONLY (tests/, smoke, bench cpu leg) -- scalar restatement of the mesh extraction of the `--watertight`
branch: sample/generate_text.py:139-141 `mcubes.marching_cubes(udf, 0.01)`.

PyMCubes (third-party, not vendored by the reference, absent here) implements the classic marching cubes of Lorensen & Cline
with the published 256 x 16 triangle table; the reference ships the same table as CASESCLASSIC
(meshudf/_marching_cubes_lewiner_luts.py), from which surfd_b200/_mc_classic_lut.py is generated.  PARITY UNPINNED: no golden
vector of mcubes exists in the reference and the package cannot be run here.  This file restates the published algorithm one
cube at a time (P. Bourke, "Polygonising a scalar field"): pattern bit i set when corner i <= iso, for every triangle of
TRI[pattern] the three edge crossings, linearly interpolated, one vertex per lattice edge.
"""
import numpy as np

from surfd_b200._mc_classic_lut import EDGE_ENDS, TRI

CORNERS = ((0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1))


def marching_cubes_loop(vol, iso):
    """(vertices float64 [V,3] in index units (axis order), faces int64 [F,3]); vertices in first-use order"""
    vol = np.asarray(vol, dtype=np.float32)
    n0, n1, n2 = vol.shape
    index, verts, faces = {}, [], []
    for i in range(n0 - 1):
        for j in range(n1 - 1):
            for k in range(n2 - 1):
                pat = 0
                for c, (a, b, d) in enumerate(CORNERS):
                    if vol[i + a, j + b, k + d] <= iso:
                        pat |= 1 << c
                row = TRI[pat]
                m = 0
                while m < 16 and row[m] >= 0:
                    tri = []
                    for e in row[m:m + 3]:
                        pa = tuple(np.add((i, j, k), EDGE_ENDS[e][0]))
                        pb = tuple(np.add((i, j, k), EDGE_ENDS[e][1]))
                        if pa > pb:
                            pa, pb = pb, pa
                        key = (pa, pb)
                        if key not in index:
                            va, vb = np.float64(vol[pa]), np.float64(vol[pb])
                            w = (iso - va) / (vb - va)
                            index[key] = len(verts)
                            verts.append(np.asarray(pa, np.float64) + w * (np.asarray(pb, np.float64) - np.asarray(pa, np.float64)))
                        tri.append(index[key])
                    faces.append(tri)
                    m += 3
    return np.asarray(verts, np.float64).reshape(-1, 3), np.asarray(faces, np.int64).reshape(-1, 3)


def canonical_faces(verts, faces):
    """order-independent form: every face as its three vertex coordinates, rotated so the smallest comes first (orientation
    kept), then the faces sorted"""
    out = []
    for f in faces:
        p = [tuple(np.round(verts[v], 9)) for v in f]
        s = min(range(3), key=lambda t: p[t])
        out.append((p[s], p[(s + 1) % 3], p[(s + 2) % 3]))
    return sorted(out)
