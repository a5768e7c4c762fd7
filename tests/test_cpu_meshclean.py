"""Post-MC mesh clean-up (meshudf.py:379-434) and output stage (generate_uncond.py:113-122): the torch implementation
(device-agnostic primitives, run here on CPU tensors with require_cuda=False) against the independent numpy / networkx
restatement in oracle/meshclean_oracle.py.  Parity with trimesh / pymeshlab themselves is UNPINNED (absent dependencies)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import meshclean_oracle as MO
from surfd_b200 import meshclean as MC, output as OUT


def _mesh(tag):
    g = np.load(os.path.join(GOLDEN, "mc_fields.npz"))
    v = g[tag + "_v"].astype(np.float64)
    f = g[tag + "_f"].reshape(-1, 3).astype(np.int64)
    return v, f


def _damage(v, f, seed):
    """holes (single triangles removed + one fan of two -> a quad hole), duplicated faces (one with rotated winding),
    degenerate faces, duplicated vertices, an unreferenced vertex"""
    rng = np.random.default_rng(seed)
    F = len(f)
    drop = rng.choice(F, size=12, replace=False)
    # a quad hole: two faces sharing an edge
    es = np.sort(f[:, [0, 1, 1, 2, 2, 0]].reshape(-1, 2), axis=1)
    key = es[:, 0] * len(v) + es[:, 1]
    order = np.argsort(key, kind="stable")
    pair = None
    for i in range(len(order) - 1):
        if key[order[i]] == key[order[i + 1]]:
            a, b = order[i] // 3, order[i + 1] // 3
            if a not in drop and b not in drop:
                pair = (a, b)
                break
    keep = np.ones(F, bool)
    keep[drop[:8]] = False
    keep[list(pair)] = False
    f2 = f[keep]
    dup = f2[rng.choice(len(f2), 5, replace=False)]
    dup[0] = dup[0][[1, 2, 0]]
    dup[1] = dup[1][::-1]
    deg = np.stack([f2[3][[0, 0, 1]], f2[9][[2, 1, 2]]])
    # duplicate vertices: re-point some faces to fresh copies of their vertices, and add an unreferenced vertex
    v2 = np.concatenate([v, v[f2[20]], np.array([[9.0, 9.0, 9.0]])])
    f2 = f2.copy()
    f2[20] = np.arange(len(v), len(v) + 3)
    f3 = np.concatenate([f2[:50], dup, f2[50:], deg])
    return v2, f3


def _free_choices(v, f):
    """faces that may legitimately differ between the two restatements: 2 per 4-cycle hole (which diagonal networkx's
    traversal picks) + 2 per small cycle through a vertex where several boundary loops touch (which cycle basis it picks)"""
    import networkx as nx
    mv, mf = MO.merge_vertices(v, f)
    mv, mf = MO.process(mv, mf)
    e = MO.faces_to_edges(mf)
    be = e[MO.group_rows_count1(np.sort(e, 1))]
    g = nx.Graph()
    g.add_edges_from(be.tolist())
    n = 0
    for c in nx.cycle_basis(g):
        if len(c) == 4 or (len(c) == 3 and any(g.degree(x) != 2 for x in c)):
            n += 2
    return n


def _face_set(f):
    f = np.asarray(f)
    rot = np.argmin(f, axis=1)
    r = np.stack([f[np.arange(len(f)), (rot + k) % 3] for k in range(3)], 1)
    return set(map(tuple, r.tolist()))


@pytest.mark.parametrize("tag,seed", [("sphere_32_0.0", 0), ("torus_48_0.3", 1), ("hemi_64_1.0", 2), ("two_40_0.0", 3)])
def test_clean_mesh_matches_oracle(tag, seed):
    v, f = _damage(*_mesh(tag), seed)
    ov, of = MO.clean_mesh(v, f)
    pv, pf = MC.clean_mesh(torch.from_numpy(v), torch.from_numpy(f), require_cuda=False)
    pv, pf = pv.numpy(), pf.numpy()
    slack = _free_choices(v, f)
    assert pv.shape == ov.shape and abs(pf.shape[0] - of.shape[0]) <= slack
    if slack == 0:
        assert np.array_equal(pv, ov)                     # same merge order, same float64 smoothing arithmetic
    else:                                                 # border smoothing sees the differently filled holes
        assert np.abs(pv - ov).max() < 1.5 and (np.abs(pv - ov).max(axis=1) > 0).mean() < 0.10
    # faces: identical up to the free choices of networkx.cycle_basis (rotation of a filled triangle, diagonal of a quad hole)
    a, b = set(map(tuple, np.sort(pf, 1).tolist())), set(map(tuple, np.sort(of, 1).tolist()))
    assert len(a - b) <= slack and len(b - a) <= slack, (len(a - b), len(b - a), slack)
    if slack == 0:
        assert _face_set(pf) == _face_set(of)             # consistently wound input: same orientation too
    n_orig = len(_face_set(f))
    assert len(a) >= n_orig - 9 - 2 + 8 - slack           # the triangle holes were filled again
    # the clean mesh has no duplicate faces and no unreferenced vertices
    assert len(np.unique(np.sort(pf, 1), axis=0)) == len(pf)
    assert set(np.unique(pf).tolist()) == set(range(len(pv)))


def test_clean_mesh_steps():
    v, f = _mesh("hemi_40_0.0")
    tv, tf = torch.from_numpy(v), torch.from_numpy(f)
    mv, mf = MC.merge_vertices(tv, tf)
    ov, of = MO.merge_vertices(v, f)
    assert np.array_equal(mv.numpy(), ov) and np.array_equal(mf.numpy(), of)
    assert np.array_equal(MC.unique_faces_mask(mf).numpy(), MO.unique_faces(of))
    assert np.array_equal(MC.nondegenerate_faces_mask(mv, mf).numpy(), MO.nondegenerate(ov, of))
    # the open hemisphere has a border: it moves, the interior does not
    sv = MC.smooth_border_vertices(mv, mf).numpy()
    assert np.array_equal(sv, MO.smooth_borders(ov, of))
    moved = np.abs(sv - ov).max(axis=1) > 0
    assert 0 < moved.sum() < len(ov) // 4
    with pytest.raises(RuntimeError):
        MC.clean_mesh(tv, tf)                               # product path: device tensors only


@pytest.mark.parametrize("tag", ["torus_48_0.3", "hemi_64_1.0", "two_40_0.0"])
def test_output_stage_matches_oracle(tag, tmp_path):
    v, f = _mesh(tag)
    v, f = MO.merge_vertices(v, f)
    # a second, small component: a copy of 60 faces of the mesh, moved away
    sub = f[100:160]
    ids = np.unique(sub)
    v = np.concatenate([v, v[ids] + 100.0])
    f = np.concatenate([f, np.searchsorted(ids, sub) + (len(v) - len(ids))])
    tv, tf = torch.from_numpy(v), torch.from_numpy(f)
    sm = OUT.laplacian_smooth(tv, tf).numpy()
    ref = MO.laplacian_smooth(v, f)
    assert np.abs(sm - ref).max() < 2e-3 * np.abs(v).max() / 48    # float32 accumulation order differs (scatter-add); cotangent weights amplify it
    lab = OUT.face_components(tf, len(v)).numpy()
    rl = MO.face_components(f, len(v))
    assert np.array_equal(lab, rl)
    sizes = np.bincount(np.unique(rl, return_inverse=True)[1])
    assert len(sizes) >= 2 and sizes.max() > 1000
    thr = 1000
    pv, pf = OUT.remove_small_components(tv, tf, thr)
    ov, of = MO.remove_small_components(v, f, thr)
    assert np.array_equal(pv.numpy(), ov) and np.array_equal(pf.numpy(), of)
    assert len(of) < len(f) and len(ov) < len(v)
    # writers / reader round trip: open3d layout, then the MeshSet surface the scripts use
    p = str(tmp_path / "0.obj")
    mesh = OUT.get_o3d_mesh_from_tensors(tv.float(), tf)
    assert OUT.io.write_triangle_mesh(p, mesh)
    head = open(p).read().split("\n")[:5]
    assert head[0] == "# Created by Open3D " and head[1] == "# object name: 0" and head[2] == "# number of points: %d" % len(v)
    assert head[4].startswith("v ")
    ms = OUT.MeshSet(device="cpu")
    ms.set_verbosity(False)
    ms.load_new_mesh(p)
    assert ms.vertices.shape == tv.shape and torch.equal(ms.faces, tf)
    assert float(((ms.vertices - tv).abs() / tv.abs().clamp(min=1)).max()) < 1e-5     # %g keeps 6 significant digits
    ms.apply_coord_laplacian_smoothing()
    ms.meshing_remove_connected_component_by_face_number(mincomponentsize=thr)
    ms.save_current_mesh(p)
    rv, rf = OUT.read_obj(p)
    assert rv.shape[0] == ms.vertices.shape[0] and torch.equal(rf, ms.faces)
