"""GPU parity (-m gpu) at BASELINE.json's sizes: the 256^3 GridFiller lattice (4 levels, far-block fills
across 3 levels) + marching cubes + face filter against the fixture the REFERENCE itself produced (tests/golden/
make_golden.py gridfiller256), and the 1000-step sampler of the default persistent engine against the reference's own
1000-step p_sample_loop (make_golden.py sampler1000).

Tolerances: fp32 decoder mode -- udf <= 1e-6, query / gradient masks identical up to knife-edge threshold ties (every
mismatch must sit within 1e-6 of its threshold, and there may be at most a handful); TF32 mode -- Jaccard reported and
bounded.  Marching cubes: bit-exact against the reference's compiled Cython (oracle/_ref) run on the SAME device-produced
lattice.  Sampler: max-abs and cosine against the reference latents, bound 1e-2."""
import hashlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from surfd_b200 import synth, unet as U
from surfd_b200.decoder import UdfDecoder
from surfd_b200.meshudf import MarchingCubes, finish_mesh

pytestmark = pytest.mark.gpu
N, L = 256, 32


def _gold():
    return np.load(os.path.join(GOLDEN, "gridfiller_poly_N256.npz"))


@pytest.fixture(scope="module")
def lattice_fp32():
    g = _gold()
    dec = UdfDecoder(synth.synth_ae_poly(L)["decoder"], L)
    dec.set_precision(0)
    dec.set_latent(torch.from_numpy(g["lat"][0]))
    udf, grads, counts = dec.lattice(N, use_fast_grid_filler=True)
    udf.clamp_(min=0)
    return dec, udf, grads, counts


def test_gridfiller_256_matches_reference(lattice_fp32):
    g = _gold()
    dec, udf, grads, counts = lattice_fp32
    n_udf_ref, n_grad_ref = int(g["level_queries"].sum()), int(g["n_grad"])
    assert g["level_queries"].tolist() == [32768, 229376, 128814, 484057]        # 4 GridFiller levels (32, 64, 128, 256)
    assert abs(counts[0] - n_udf_ref) <= 8 and abs(counts[1] - n_grad_ref) <= 8, (counts, n_udf_ref, n_grad_ref)
    u = udf.reshape(-1).cpu().numpy()
    gr = grads.reshape(-1, 3).cpu().numpy()
    # values: near-surface sample (every one a gradient point of the reference) and a uniform sample (mostly far-block copies)
    assert np.abs(u[g["idx_near"]] - g["udf_near"]).max() < 1e-6
    diff_any = np.abs(u[g["idx_any"]] - g["udf_any"])
    assert (diff_any < 1e-6).mean() > 0.9999 and np.quantile(diff_any, 0.9999) < 1e-6   # a far/close tie moves a whole block's fill
    # gradient mask: identical up to threshold ties
    thr = np.float32(2.5 * 2.0 / N)
    ref_mask = np.unpackbits(g["gradmask"])[:N ** 3].astype(bool)
    mine = np.abs(gr).sum(-1) > 0
    bad = np.nonzero(ref_mask != mine)[0]
    assert len(bad) <= 8, f"{len(bad)} gradient-mask mismatches"
    assert np.all(np.abs(u[bad] - thr) < 1e-6), "a mismatch that is not a threshold tie"
    print(f"N=256 fp32: udf max err {np.abs(u[g['idx_near']] - g['udf_near']).max():.2e}, gradient-mask ties {len(bad)} of {int(ref_mask.sum())}")
    # the coarse structure of the filled lattice (which blocks carry copied far values)
    far = np.unpackbits(g["far_value_mask"])[:N ** 3].astype(bool)
    assert ((u >= np.float32(0.0199)) != far).sum() <= 64 * 8
    # gradients where both have one
    sel = ref_mask[g["idx_near"]] & mine[g["idx_near"]]
    gerr = np.abs(gr[g["idx_near"]][sel] - g["grads_near"][sel]).max(-1)
    assert np.quantile(gerr, 0.995) < 2e-4 and gerr.max() < 0.1


def test_marching_cubes_256_on_decoder_field_is_bit_exact(lattice_fp32, ref_mc):
    g = _gold()
    dec, udf, grads, counts = lattice_fp32
    mc = MarchingCubes()
    v, f = mc.run_raw(udf, grads)
    nv_ref, nf_ref = (int(x) for x in g["mc_nv_nf"])
    # the reference mesh of the REFERENCE lattice: counts agree up to threshold ties of the lattice (values differ by <= 1e-6)
    assert abs(v.shape[0] - nv_ref) <= 0.001 * nv_ref and abs(f.shape[0] - nf_ref) <= 0.001 * nf_ref, (v.shape, f.shape, nv_ref, nf_ref)
    if ref_mc is None:
        pytest.skip("oracle/_ref not built: bit-exactness against the compiled reference cannot be checked")
    rv, rf = ref_mc(udf.cpu().numpy(), grads.cpu().numpy())
    assert rv.shape[0] == v.shape[0] and rf.shape[0] == f.numel()
    assert np.array_equal(rv, v.cpu().numpy()) and np.array_equal(rf.reshape(-1, 3), f.cpu().numpy())
    vertices, faces = finish_mesh(v, f, N)
    if v.shape[0] == nv_ref and f.shape[0] == nf_ref:
        sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()   # noqa: E731
        same = sha(vertices.cpu().numpy()) == str(g["mc_verts_sha"]) and sha(faces.cpu().numpy()) == str(g["mc_faces_sha"])
        print("N=256 mesh identical to the reference mesh of the reference lattice (SHA-256):", same)
        head = np.abs(vertices.cpu().numpy()[:4096] - g["mc_verts_head"]).max()
        assert head < 1e-5 and np.array_equal(faces.cpu().numpy()[:8192], g["mc_faces_head"])
    # face filter: the reference keeps every face of this mesh
    keep = dec.face_filter(vertices, faces, N)
    assert abs(int(keep.sum()) - int(g["n_keep"])) <= 0.001 * nf_ref


def test_tf32_lattice_256_jaccard(lattice_fp32):
    g = _gold()
    dec, udf, grads, counts = lattice_fp32
    dec.set_precision(1)
    try:
        u2, g2, c2 = dec.lattice(N, use_fast_grid_filler=True)
    finally:
        dec.set_precision(0)
    m1 = (grads.abs().sum(-1) > 0).reshape(-1)
    m2 = (g2.abs().sum(-1) > 0).reshape(-1)
    jac = float((m1 & m2).sum()) / float((m1 | m2).sum())
    err = float((u2.clamp(min=0) - udf)[m1.reshape(udf.shape) & m2.reshape(udf.shape)].abs().max())
    print(f"N=256 TF32 vs fp32: gradient-mask Jaccard {jac:.5f}, udf max err on the shared band {err:.2e}, queries {c2} vs {counts}")
    assert jac > 0.99 and err < 5e-4   # reported; the 2e-4 of DESIGN.md 5 is the bound on the golden query points


CASES_1000 = (("uncond32_b8", 32, "no_cond", 8, 1.0), ("text64_cfg_b4", 64, "img", 4, 4.0))


@pytest.mark.parametrize("case", CASES_1000, ids=lambda c: c[0])
def test_thousand_step_latents_match_reference(case):
    """C2/C3's and C5's sampler, all 1000 steps, default engine (persistent kernel, tcgen05 units): the deviation that
    SURVEY.md 8(d) asks to be reported -- max-abs and cosine against the reference's own p_sample_loop output."""
    tag, L_, cond, B, guidance = case
    path = os.path.join(GOLDEN, "sampler1000.npz")
    if not os.path.exists(path):
        pytest.skip("sampler1000.npz fixture missing")
    g = np.load(path)
    net = U.UNetSampler(synth.synth_mdm(L_, cond), L_, cond, max_batch=8)
    S = U.SpacedSchedule(U.cosine_betas(), U.space_timesteps(1000, [1000]))
    noise = torch.randn(1001, B, L_, generator=torch.Generator().manual_seed(10))
    ctx = torch.from_numpy(g[tag + "_ctx"]) if cond == "img" else None
    out = net.sample(S, noise, ctx, None, guidance)
    torch.cuda.synchronize(); net.status()
    ref = torch.from_numpy(g[tag + "_sample"]).reshape(B, L_)
    o = out.reshape(B, L_).cpu()
    err = float((o - ref).abs().max())
    cos = float(torch.nn.functional.cosine_similarity(o.reshape(1, -1), ref.reshape(1, -1)))
    print(f"{tag}: 1000-step latents vs reference: max-abs {err:.3e}, cosine {cos:.8f}, |x0| max {float(ref.abs().max()):.3f}")
    assert torch.isfinite(o).all() and err < 1e-2 and cos > 0.9999
