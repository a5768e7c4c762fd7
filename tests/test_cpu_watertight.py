"""`--watertight` extraction (surfd_b200/watertight.py; generate_text.py:132-158) against the scalar oracle and through the
properties a closed iso-surface has.  torch on CPU -- the same code runs on the device in the CLI."""
import math

import numpy as np
import pytest
import torch

from oracle.watertight_oracle import canonical_faces, marching_cubes_loop
from surfd_b200.watertight import marching_cubes, watertight_mesh


def _shell(N, r=0.6, sx=1.0, sy=1.0, sz=1.0):
    g = torch.linspace(-1, 1, N)
    a, b, c = torch.meshgrid(g, g, g, indexing="ij")
    return ((sx * a * a + sy * b * b + sz * c * c).sqrt() - r).abs()


def _edge_counts(f, nv):
    e = torch.cat([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
    und = e.min(1).values * nv + e.max(1).values
    _, c = torch.unique(und, return_counts=True)
    directed = e[:, 0] * nv + e[:, 1]
    return c, torch.unique(directed).numel() == directed.numel()


@pytest.mark.parametrize("seed", [0, 1])
def test_matches_scalar_oracle_on_random_smooth_fields(seed):
    g = torch.Generator().manual_seed(seed)
    coarse = torch.rand(1, 1, 5, 6, 4, generator=g)
    vol = torch.nn.functional.interpolate(coarse, size=(14, 17, 11), mode="trilinear", align_corners=True)[0, 0]
    iso = 0.5
    v, f = marching_cubes(vol, iso)
    ov, of = marching_cubes_loop(vol.numpy(), iso)
    assert v.shape[0] == ov.shape[0] and f.shape[0] == of.shape[0] and f.shape[0] > 100
    assert canonical_faces(v.numpy(), f.numpy()) == canonical_faces(ov, of)


def test_exact_ties_and_empty_volume():
    vol = torch.zeros(6, 6, 6)
    vol[2:4, 2:4, 2:4] = 1.0
    v, f = marching_cubes(vol, 0.0)          # lattice values equal to the level count as "inside" (<=): nothing crosses at 0 from below
    ov, of = marching_cubes_loop(vol.numpy(), 0.0)
    assert canonical_faces(v.numpy(), f.numpy()) == canonical_faces(ov, of)
    v, f = marching_cubes(torch.full((5, 5, 5), 0.3), 0.01)
    assert v.shape == (0, 3) and f.shape == (0, 3)
    with pytest.raises(RuntimeError, match="No surface"):
        watertight_mesh(torch.full((5, 5, 5), 0.3))


def test_shell_of_an_unsigned_field_is_closed_and_oriented():
    N, iso = 72, 0.04
    v, f = marching_cubes(_shell(N, 0.6, 1.0, 1.3, 0.8), iso)
    c, consistent = _edge_counts(f, v.shape[0])
    assert int(c.min()) == 2 and int(c.max()) == 2          # every edge shared by exactly two faces: watertight
    assert consistent                                         # ... with opposite directions: consistently oriented
    n_edges = c.numel()
    assert v.shape[0] - n_edges + f.shape[0] == 4             # two nested spheres (Euler characteristic 2 each)
    # vertices lie on the level set of the trilinear interpolant: check against the analytic field
    w = v / (N - 1) * 2 - 1
    val = ((w[:, 0] ** 2 + 1.3 * w[:, 1] ** 2 + 0.8 * w[:, 2] ** 2).sqrt() - 0.6).abs()
    assert float((val - iso).abs().max()) < 2e-3
    # enclosed volume of the shell, sign: normals point towards lower values like mcubes (inside = above the level)
    p = w[f]
    vol = float((p[:, 0] * torch.cross(p[:, 1], p[:, 2], dim=1)).sum() / 6)
    want = 4 / 3 * math.pi * ((0.6 + iso) ** 3 - (0.6 - iso) ** 3) / math.sqrt(1.3 * 0.8)
    assert abs(-vol - want) / want < 0.02


def test_small_components_are_removed_and_negative_values_clamped():
    N = 64
    big = _shell(N, 0.6)
    g = torch.linspace(-1, 1, N)
    a, b, c = torch.meshgrid(g, g, g, indexing="ij")
    small = (((a - 0.85) ** 2 + (b - 0.85) ** 2 + (c - 0.85) ** 2).sqrt() - 0.06).abs()
    udf = torch.minimum(big, small)
    udf[0, 0, 0] = -1.0                                       # udf[udf < 0] = 0 (generate_text.py:136)
    v_all, f_all = marching_cubes(udf.clamp(min=0), 0.03)
    v, f = watertight_mesh(udf, iso=0.03, mincomponentsize=5000)
    assert 0 < f.shape[0] < f_all.shape[0] and v.shape[0] < v_all.shape[0]
    w = v / (N - 1) * 2 - 1
    assert float(w.norm(dim=1).max()) < 0.7                   # the blob at the corner (and the clamped corner point) are gone
    c, consistent = _edge_counts(f, v.shape[0])
    assert int(c.min()) == 2 and int(c.max()) == 2 and consistent
    assert v.dtype == torch.float32 and f.dtype == torch.int64 and int(f.max()) == v.shape[0] - 1


@pytest.mark.parametrize("seed", range(4))
def test_random_fields_give_closed_oriented_surfaces_away_from_the_lattice_boundary(seed):
    """the 256-entry table is consistent on ambiguous faces: whatever the field, every interior edge of the extracted surface is
    shared by exactly two faces running in opposite directions (edges on the lattice boundary may be open)"""
    g = torch.Generator().manual_seed(100 + seed)
    vol = torch.nn.functional.interpolate(torch.rand(1, 1, 7, 7, 7, generator=g), size=(24, 24, 24), mode="trilinear", align_corners=True)[0, 0]
    vol = vol + 0.15 * torch.rand(24, 24, 24, generator=g)                 # noise: many ambiguous cubes
    v, f = marching_cubes(vol, 0.55)
    assert f.shape[0] > 500
    e = torch.cat([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
    nv = v.shape[0]
    on_boundary = ((v <= 0) | (v >= 23)).any(1)
    und = e.min(1).values * nv + e.max(1).values
    uniq, inv, cnt = torch.unique(und, return_inverse=True, return_counts=True)
    interior = ~(on_boundary[e[:, 0]] & on_boundary[e[:, 1]])
    assert bool((cnt[inv][interior] == 2).all())
    directed = e[:, 0] * nv + e[:, 1]
    assert torch.unique(directed).numel() == directed.numel()
    ov, of = marching_cubes_loop(vol.numpy(), 0.55)
    assert canonical_faces(v.numpy(), f.numpy()) == canonical_faces(ov, of)
