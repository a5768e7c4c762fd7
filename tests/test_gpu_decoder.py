"""GPU parity (-m gpu): the CUDA decoder / lattice / face filter through the C-ABI against the reference-generated
goldens and the numpy oracle.  Tolerances (fp32 mode): udf <= 1e-6 abs on values in [0,0.1]; gradient
components <= 2e-4 (unit vectors, ~1e-3 rad) with an identical zero set; query masks identical up to
knife-edge threshold ties (reported, bounded)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import decoder_oracle as O
from surfd_b200 import synth
from surfd_b200.decoder import UdfDecoder

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("L", [32, 64])
def test_query_matches_reference_golden(L):
    g = np.load(os.path.join(GOLDEN, f"decoder_L{L}.npz"))
    dec = UdfDecoder(synth.synth_ae_rand(L, 4321)["decoder"], L)
    dec.set_latent(torch.from_numpy(g["lat"][0]))
    udf, grads = dec.query(torch.from_numpy(g["pts"]), want_grad=True)
    udf, grads = udf.cpu().numpy(), grads.cpu().numpy()
    assert np.abs(udf - g["udf"]).max() < 1e-6
    # the field is piecewise linear: a pre-activation within rounding of 0 flips a ReLU mask and moves the gradient by a
    # finite amount, so a handful of points may differ visibly; everything else agrees to ~1e-3 rad.
    gerr = np.abs(grads - g["grads"]).max(-1)
    assert np.quantile(gerr, 0.995) < 2e-4 and gerr.max() < 0.1, (np.quantile(gerr, 0.995), gerr.max())
    assert ((np.abs(grads).sum(-1) == 0) == (np.abs(g["grads"]).sum(-1) == 0)).all()
    only = dec.query(torch.from_numpy(g["pts"]))
    assert torch.equal(only.cpu(), torch.from_numpy(udf))     # forward-only path == forward of the gradient path


def test_query_ragged_sizes_and_chunking_are_consistent():
    L = 32
    sd = synth.synth_ae_rand(L, 4321)["decoder"]
    gen = torch.Generator().manual_seed(3)
    lat = torch.randn(L, generator=gen)
    pts = torch.rand(5000, 3, generator=gen) * 2 - 1
    big = UdfDecoder(sd, L)
    small = UdfDecoder(sd, L, max_chunk_points=1024)          # forces 5 chunks with a ragged tail
    big.set_latent(lat); small.set_latent(lat)
    u0, g0 = big.query(pts, want_grad=True)
    u1, g1 = small.query(pts, want_grad=True)
    assert torch.equal(u0, u1) and torch.equal(g0, g1)        # per-point results do not depend on batching
    for m in (0, 1, 7, 129):
        u = big.query(pts[:m])
        assert u.shape == (m,) and torch.equal(u, u0[:m])
    ref = O.forward(sd, lat.numpy(), pts.numpy()[:512])
    assert np.abs(u0[:512].cpu().numpy() - ref).max() < 1e-6


def test_poly_decoder_field_is_the_polytope_udf():
    L = 32
    dec = UdfDecoder(synth.synth_ae_poly(L)["decoder"], L)
    gen = torch.Generator().manual_seed(5)
    lat = torch.randn(L, generator=gen)
    pts = torch.rand(20000, 3, generator=gen) * 2 - 1
    dec.set_latent(lat)
    udf, grads = dec.query(pts, want_grad=True)
    exact, m = synth.poly_udf(pts, lat)
    assert (udf.cpu().double() - exact).abs().max() < 1e-6
    # gradient = -grad(udf) = -sign(m) * n_argmax near the surface (unit normal of the nearest face)
    n = synth.poly_planes(32).double()
    r = synth.poly_offsets(lat).double()
    a = pts.double() @ n.T - r
    top2 = a.topk(2, dim=-1).values
    sel = (m.abs() < 0.04) & ((top2[:, 0] - top2[:, 1]) > 1e-3)     # away from polytope edges
    expect = -(torch.sign(m)[:, None] * n[a.argmax(-1)])
    assert (grads.cpu().double()[sel] - expect[sel]).abs().max() < 1e-4


@pytest.mark.parametrize("mode", ["gf", "dense"])
def test_lattice_matches_reference_golden(mode):
    g = np.load(os.path.join(GOLDEN, "gridfiller_poly_N64.npz"))
    N, L = 64, 32
    dec = UdfDecoder(synth.synth_ae_poly(L)["decoder"], L)
    dec.set_latent(torch.from_numpy(g["lat"][0]))
    udf, grads, counts = dec.lattice(N, use_fast_grid_filler=(mode == "gf"))
    udf, grads = udf.cpu().numpy(), grads.cpu().numpy()
    assert np.abs(udf - g[mode + "_udf"]).max() < 1e-6
    mask = np.unpackbits(g[mode + "_gradmask"])[:N ** 3].astype(bool).reshape(N, N, N)
    mine = np.abs(grads).sum(-1) > 0
    ties = int((mask != mine).sum())
    assert ties <= 4, f"{ties} gradient-mask mismatches (only threshold ties are tolerated)"
    both = mask & mine
    ref_g = np.zeros((N, N, N, 3), np.float32); ref_g[mask] = g[mode + "_grads"].astype(np.float32)
    assert np.abs(grads[both] - ref_g[both]).max() < 1e-3      # golden grads stored as fp16
    n_udf_ref = int(g[mode + "_calls"][0]) - int(g[mode + "_calls"][1])
    assert abs(counts[0] - n_udf_ref) <= 4 and abs(counts[1] - int(g[mode + "_calls"][1])) <= 4


def test_gridfiller_prunes_and_agrees_with_dense_near_the_surface():
    N, L = 128, 32
    dec = UdfDecoder(synth.synth_ae_poly(L)["decoder"], L)
    gen = torch.Generator().manual_seed(11)
    dec.set_latent(torch.randn(L, generator=gen))
    u_gf, g_gf, c_gf = dec.lattice(N, True)
    u_d, g_d, c_d = dec.lattice(N, False)
    assert c_d[0] == N ** 3 and c_gf[0] < 0.6 * N ** 3         # coarse-to-fine skips far blocks
    near = u_d < 1.5 * 1.7 * (2.0 / 64) * 0.5                  # well inside the finest "close" band
    assert torch.equal(u_gf[near], u_d[near])                  # same points evaluated -> identical values
    thr = np.float32(2.5 * 2.0 / N)
    sel = u_d < thr
    assert torch.equal(g_gf[sel], g_d[sel])


def test_face_filter_matches_oracle():
    L, N = 32, 64
    sd = synth.synth_ae_poly(L)["decoder"]
    dec = UdfDecoder(sd, L)
    gen = torch.Generator().manual_seed(2)
    lat = torch.randn(L, generator=gen)
    dec.set_latent(lat)
    # random small triangles near the surface r ~ 0.5
    c = torch.nn.functional.normalize(torch.randn(300, 3, generator=gen, dtype=torch.float64), dim=-1) * (0.5 + 0.04 * torch.randn(300, 1, generator=gen, dtype=torch.float64))
    verts = (c[:, None, :] + 0.02 * torch.randn(300, 3, 3, generator=gen, dtype=torch.float64)).reshape(-1, 3)
    faces = torch.arange(900, dtype=torch.int32).reshape(300, 3)
    keep = dec.face_filter(verts, faces, N).cpu().numpy().astype(bool)
    v = verts.numpy(); f = faces.numpy()
    e0 = f[:, [0, 1, 2]].reshape(-1); e1 = f[:, [1, 2, 0]].reshape(-1)
    pts = np.vstack([v[e0], v[e1], (v[e0] + v[e1]) / 2]).astype(np.float32)        # meshudf.py:358-367
    udf = O.forward(sd, lat.numpy(), pts)
    face_idx = np.hstack([np.repeat(np.arange(300), 3)] * 3)
    bad = np.unique(face_idx[udf > np.float32(1.0 / N)])
    ref_keep = np.ones(300, bool); ref_keep[bad] = False
    margin = np.abs(udf - 1.0 / N).reshape(3, 300, 3).min(axis=(0, 2)) > 1e-6   # skip knife-edge faces
    assert (keep[margin] == ref_keep[margin]).all()
    assert 0 < keep.sum() < 300


@pytest.mark.parametrize("precision", [0, 1], ids=["fp32", "tf32"])
def test_face_filter_vertex_sharing_equals_nine_queries_per_face(precision):
    """the filter evaluates every vertex once + 3 midpoints per face; the decisions must equal the reference's literal
    9 queries per face (same decoder, same precision: a point's udf does not depend on where it sits in a batch)"""
    from surfd_b200.meshudf import MarchingCubes, finish_mesh
    L, N = 32, 64
    dec = UdfDecoder(synth.synth_ae_poly(L)["decoder"], L)
    dec.set_precision(precision)
    dec.set_latent(torch.randn(L, generator=torch.Generator().manual_seed(5)) * 0.7)
    udf, grads, _ = dec.lattice(N, True)
    v, f = MarchingCubes().run_raw(udf.clamp(min=0), grads)
    verts, faces = finish_mesh(v, f, N)                       # float64 [V,3], shared vertices
    assert faces.shape[0] > 1000 and verts.shape[0] < faces.shape[0]
    keep = dec.face_filter(verts, faces, N).bool()
    fl = faces.long()
    e0 = fl[:, [0, 1, 2]].reshape(-1); e1 = fl[:, [1, 2, 0]].reshape(-1)
    pts = torch.cat([verts[e0], verts[e1], (verts[e0] + verts[e1]) / 2]).float()      # meshudf.py:358-367
    far = dec.query(pts) > (1.0 / N)
    bad = far.reshape(3, faces.shape[0], 3).any(dim=2).any(dim=0)
    assert torch.equal(keep, ~bad)
    assert 0 < int(keep.sum())


def test_missing_latent_and_foreign_callable_fail_loudly():
    from surfd_b200.meshudf import get_mesh_from_udf
    dec = UdfDecoder(synth.synth_ae_rand(32, 1)["decoder"], 32)
    with pytest.raises(ValueError):
        dec.query(torch.zeros(4, 3))
    with pytest.raises(TypeError):
        get_mesh_from_udf(lambda c: c[:, 0], (-1, 1), 0.1, N=32, differentiable=False)
