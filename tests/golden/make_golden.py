#!/usr/bin/env python
"""Generate the committed golden vectors by running the REFERENCE itself (imported from /root/reference,
CPU fp32) on seeded inputs.  Run in the build container only:  python tests/golden/make_golden.py [what...]

Fixtures (tests/golden/*.npz) are small on purpose; the synthetic checkpoints are regenerated from seeds
by surfd_b200.synth, so only inputs that are not seed-derivable and the reference outputs are stored.
"""
import os, sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"
sys.path[:0] = [os.path.join(ROOT, "oracle", "stubs"), os.path.join(ROOT, "oracle", "_ref"), REF]
sys.dont_write_bytecode = True
OUT = os.path.dirname(os.path.abspath(__file__))
torch.Tensor.cuda = lambda self, *a, **k: self          # the reference hard-codes .cuda()
torch.nn.Module.cuda = lambda self, *a, **k: self


def ref_decoder(L, seed):
    from AutoEncoder.models.cbndec import CbnDecoder
    from AutoEncoder.models.coordsenc import CoordsEncoder
    from surfd_b200.synth import synth_ae_rand
    ck = synth_ae_rand(L, seed)
    dec = CbnDecoder(63, L, 512, 5)
    dec.load_state_dict(ck["decoder"], strict=True)
    dec.eval()
    for p in dec.parameters():
        p.requires_grad = False
    return dec, CoordsEncoder(), ck


def golden_decoder():
    from meshudf.meshudf import sample_udf, sample_grads
    for L in (32, 64):
        dec, enc, ck = ref_decoder(L, 4321)
        g = torch.Generator().manual_seed(100 + L)
        lat = 0.7 * torch.randn(1, L, generator=g)
        pts = torch.rand(2048, 3, generator=g) * 2 - 1
        # also a run of exact lattice coordinates (N=64) as the reference builds them
        idx = torch.arange(0, 1024)
        N = 64
        vs = 2.0 / (N - 1)
        lat_pts = torch.stack([(idx // 16 % 8).float() * vs + (-1), (idx // 4 % 4 * 5).float() * vs + (-1),
                               (idx % 64).float() * vs + (-1)], -1)
        pts = torch.cat([pts, lat_pts], 0)

        def udf_func(c):
            c = enc.encode(c.unsqueeze(0))
            p = dec(c, lat).squeeze(0)
            p = torch.sigmoid(p)
            return (1 - p) * 0.1
        udf = sample_udf(udf_func, pts, 2 ** 16)
        grads = sample_grads(udf_func, pts, 2 ** 16)
        np.savez_compressed(os.path.join(OUT, f"decoder_L{L}.npz"), lat=lat.numpy(), pts=pts.numpy(),
                            udf=udf.numpy(), grads=grads.numpy())
        print("decoder golden L", L, "udf range", float(udf.min()), float(udf.max()))


def ref_poly(L):
    from AutoEncoder.models.cbndec import CbnDecoder
    from AutoEncoder.models.coordsenc import CoordsEncoder
    from surfd_b200.synth import synth_ae_poly
    ck = synth_ae_poly(L)
    dec = CbnDecoder(63, L, 512, 5)
    dec.load_state_dict(ck["decoder"], strict=True)
    dec.eval()
    for p in dec.parameters():
        p.requires_grad = False
    return dec, CoordsEncoder(), ck


def golden_gridfiller():
    """Reference GridFiller / dense lattice + reference marching cubes on the 'poly' decoder, N=64."""
    from meshudf.meshudf import GridFiller, get_udf_and_grads
    from meshudf._marching_cubes_lewiner import udf_mc_lewiner
    L, N = 32, 64
    dec, enc, ck = ref_poly(L)
    g = torch.Generator().manual_seed(7)
    lat = torch.randn(1, L, generator=g)
    calls = {"n": 0}

    def udf_func(c):
        calls["n"] += c.shape[0]
        c = enc.encode(c.unsqueeze(0))
        p = dec(c, lat).squeeze(0)
        p = torch.sigmoid(p)
        return (1 - p) * 0.1
    out = {"lat": lat.numpy()}
    for mode in ("gf", "dense"):
        calls["n"] = 0
        if mode == "gf":
            udf, grads = GridFiller(N).fill_grid(udf_func, 2 ** 16)
        else:
            udf, grads = get_udf_and_grads(udf_func, (-1, 1), 0.1, N, 2 ** 16)
        udf = udf.clone(); udf[udf < 0] = 0
        u, gr = udf.detach().numpy(), grads.detach().numpy()
        v, f, _, _ = udf_mc_lewiner(u, gr, spacing=[2.0 / (N - 1)] * 3)
        n_grad = int((np.abs(gr).sum(-1) > 0).sum())
        print(mode, "calls", calls["n"], "grad pts", n_grad, "V", v.shape, "F", f.shape)
        out[mode + "_udf"] = u.astype(np.float32)
        out[mode + "_gradmask"] = np.packbits(np.abs(gr).sum(-1) > 0)
        sel = np.abs(gr).sum(-1) > 0
        out[mode + "_grads"] = gr[sel].astype(np.float16)
        out[mode + "_calls"] = np.array([calls["n"], n_grad])
        out[mode + "_nv_nf"] = np.array([v.shape[0], f.shape[0]])
    out["gf_udf"] = out["gf_udf"].astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "gridfiller_poly_N64.npz"), **out)


def golden_gridfiller256():
    """BASELINE size: the reference's GridFiller (4 levels: 32, 64, 128, 256) + gradients + marching cubes + UDF face filter at
    N = 256 on the 'poly' decoder.  The 67 MB + 201 MB lattices are not stored: the fixture keeps the per-level query counts,
    the query / gradient masks (bit-packed), udf / gradient samples, and counts + SHA-256 of the reference mesh."""
    import hashlib
    from meshudf import meshudf as M
    from meshudf._marching_cubes_lewiner import udf_mc_lewiner
    L, N = 32, 256
    dec, enc, ck = ref_poly(L)
    g = torch.Generator().manual_seed(7)
    lat = torch.randn(1, L, generator=g)
    calls = []

    def udf_func(c):
        calls.append(int(c.shape[0]))
        c = enc.encode(c.unsqueeze(0))
        p = dec(c, lat).squeeze(0)
        p = torch.sigmoid(p)
        return (1 - p) * 0.1
    torch.set_num_threads(os.cpu_count())
    gf = M.GridFiller(N)
    # per-level query counts: fill_grid calls sample_udf once per level (batches of 2**16), then sample_grads
    level_marks = []
    orig_sample_udf = M.sample_udf

    def counting_sample_udf(f, pts, max_batch):
        level_marks.append(int(pts.shape[0]))
        return orig_sample_udf(f, pts, max_batch)
    M.sample_udf = counting_sample_udf
    try:
        udf, grads = gf.fill_grid(udf_func, 2 ** 16)
    finally:
        M.sample_udf = orig_sample_udf
    udf = udf.clone(); udf[udf < 0] = 0
    u, gr = np.ascontiguousarray(udf.detach().numpy(), dtype=np.float32), np.ascontiguousarray(grads.detach().numpy(), dtype=np.float32)
    gmask = np.abs(gr).sum(-1) > 0
    n_grad = int(gmask.sum())
    print("levels", level_marks, "n_grad", n_grad)
    v, f, _, _ = udf_mc_lewiner(u, gr, spacing=[2.0 / (N - 1)] * 3)
    v = v + (-1)
    print("MC", v.shape, f.shape)
    # face filter, literally meshudf.py:356-379 (trimesh-free: edges / edges_face restated as in SURVEY 8(c))
    edges = f[:, [0, 1, 1, 2, 2, 0]].reshape(-1, 2)
    face_idxs = np.repeat(np.arange(len(f)), 3)
    edge_pts = v[edges]                                           # [3F, 2, 3] float64
    mid = edge_pts.mean(axis=1)
    pts = np.concatenate([edge_pts[:, 0], edge_pts[:, 1], mid], axis=0)
    with torch.no_grad():
        uf = M.sample_udf(udf_func, torch.from_numpy(pts).float(), 2 ** 16).numpy()
    uf = uf.reshape(3, -1)
    mask = (uf > (1 / N)).any(axis=0)
    remove = np.unique(face_idxs[mask])
    keep = np.ones(len(f), bool); keep[remove] = False
    print("face filter keeps", int(keep.sum()), "of", len(f))
    rng = np.random.default_rng(256)
    flat_u = u.reshape(-1)
    near = np.nonzero(gmask.reshape(-1))[0]
    idx_near = np.sort(rng.choice(near, size=min(32768, len(near)), replace=False))
    idx_any = np.sort(rng.choice(N ** 3, size=32768, replace=False))
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()   # noqa: E731
    # which lattice points were queried at all: active at the finest level they belong to == not copied from a far block.
    # Reconstructed from the reference's own masks: a point is "queried" iff it belongs to samples evaluated at some level.
    out = dict(lat=lat.numpy(), level_queries=np.array(level_marks), n_grad=np.array(n_grad),
               gradmask=np.packbits(gmask), udf_sha=np.array(sha(u)), grads_sha=np.array(sha(gr)),
               idx_near=idx_near.astype(np.int32), udf_near=flat_u[idx_near], grads_near=gr.reshape(-1, 3)[idx_near],
               idx_any=idx_any.astype(np.int32), udf_any=flat_u[idx_any],
               far_value_mask=np.packbits(flat_u >= np.float32(0.0199)),      # coarse structure of the filled lattice
               mc_nv_nf=np.array([v.shape[0], f.shape[0]]), mc_verts_sha=np.array(sha(v)), mc_faces_sha=np.array(sha(f)),
               mc_verts_head=v[:4096].astype(np.float64), mc_faces_head=f[:8192].astype(np.int32),
               keep_bits=np.packbits(keep), n_keep=np.array(int(keep.sum())))
    np.savez_compressed(os.path.join(OUT, "gridfiller_poly_N256.npz"), **out)
    print("wrote gridfiller_poly_N256.npz", os.path.getsize(os.path.join(OUT, "gridfiller_poly_N256.npz")) >> 10, "KiB")


def golden_sampler1000():
    """The reference's full 1000-step p_sample_loop (SpacedDiffusion(space_timesteps(1000,[1000]))) with injected noise:
    C2/C3's sampler (uncond, L = 32, B = 8) and C5's (context + CFG wrapper scale 4.0, L = 64, B = 4).  Final latents only."""
    import argparse
    from utils.model_util import create_model_and_diffusion, load_model_wo_clip
    from models.cfg_sampler import ClassifierFreeSampleModel
    from surfd_b200.synth import synth_mdm
    torch.set_num_threads(os.cpu_count())
    out = {}
    for tag, L, cond, B in (("uncond32_b8", 32, "no_cond", 8), ("text64_cfg_b4", 64, "img", 4)):
        args = argparse.Namespace(cond_mode=cond, num_actions=9, arch="OpenUNet", dataset="x", noise_schedule="cosine",
                                  sigma_small=True, clip_value=0.1)
        model, diff = create_model_and_diffusion(args)
        load_model_wo_clip(model, synth_mdm(L, cond))
        model.eval()
        g = torch.Generator().manual_seed(10)
        noise = torch.randn(1001, B, L, generator=g)
        draws = [noise[1 + k][:, None, :].clone() for k in range(1000)]
        orig = torch.randn_like
        torch.randn_like = lambda ref, *a, **k: draws.pop(0)
        mk = {"y": {}}
        m = model
        if cond == "img":
            ctx = 0.5 * torch.randn(B, 512, generator=torch.Generator().manual_seed(77))
            mk = {"y": {"context": ctx, "scale": torch.ones(B) * 4.0}}

            class _Mode(str):
                def __contains__(self, item):
                    return item == "img"
            model.cond_mode = _Mode("text")
            m = ClassifierFreeSampleModel(model)
            out[tag + "_ctx"] = ctx.numpy()
        try:
            with torch.no_grad():
                res = diff.p_sample_loop(m, (B, 1, L), noise=noise[0][:, None, :].clone(), clip_denoised=False, model_kwargs=mk,
                                         skip_timesteps=0, init_image=None, progress=False, dump_steps=None, const_noise=False)
        finally:
            torch.randn_like = orig
        out[tag + "_sample"] = res.numpy()
        print(tag, "1000-step sample absmax", float(res.abs().max()))
    np.savez_compressed(os.path.join(OUT, "sampler1000.npz"), **out)


def golden_mc():
    """Reference Cython marching_cubes_udf (compiled into oracle/_ref) on the analytic fields of tests/fields.py."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from fields import analytic_field, MC_CASES
    from meshudf import _marching_cubes_lewiner_cy as cy
    from meshudf._marching_cubes_lewiner import _get_mc_luts
    L = _get_mc_luts()
    out = {}
    for kind, N, noise in MC_CASES:
        udf, g = analytic_field(kind, N, noise, seed=N)
        v, f, _, _ = cy.marching_cubes_udf(udf, g, L, 1, 0, None)
        key = f"{kind}_{N}_{noise}"
        out[key + "_v"] = v
        out[key + "_f"] = f.astype(np.int32)
        print(key, v.shape, f.shape)
    np.savez_compressed(os.path.join(OUT, "mc_fields.npz"), **out)


def golden_unet():
    """Reference MDM / SpacedDiffusion on the synthetic diffusion checkpoint: teacher-forced forwards and a 10-step
    respaced p_sample_loop with injected noise (torch.randn patched to replay the recorded draws)."""
    import argparse
    from utils.model_util import create_model_and_diffusion, load_model_wo_clip
    from diffusion.respace import SpacedDiffusion, space_timesteps
    from diffusion import gaussian_diffusion as gd
    from models.cfg_sampler import ClassifierFreeSampleModel
    from surfd_b200.synth import synth_mdm
    out = {}
    for tag, L, cond in (("uncond32", 32, "no_cond"), ("img64", 64, "img"), ("cat32", 32, "category")):
        args = argparse.Namespace(cond_mode=cond, num_actions=9, arch="OpenUNet", dataset="x", noise_schedule="cosine",
                                  sigma_small=True, clip_value=0.1)
        model, _ = create_model_and_diffusion(args)
        load_model_wo_clip(model, synth_mdm(L, cond))
        model.eval()
        g = torch.Generator().manual_seed({"uncond32": 11, "img64": 12, "cat32": 13}[tag])
        B = 3
        x = torch.randn(B, 1, L, generator=g)
        t = torch.tensor([0, 500, 999])
        ctx = 0.5 * torch.randn(B, 512, generator=g) if cond == "img" else None
        lab = torch.tensor([0, 4, 8]) if cond == "category" else None
        y = {"context": ctx} if ctx is not None else ({"action_text": lab} if lab is not None else {})
        with torch.no_grad():
            o = model(x, t, y=y)
        out[tag + "_x"] = x.numpy(); out[tag + "_t"] = t.numpy(); out[tag + "_out"] = o.numpy()
        if ctx is not None: out[tag + "_ctx"] = ctx.numpy()
        if lab is not None: out[tag + "_lab"] = lab.numpy()
        print(tag, "forward out absmax", float(o.abs().max()))
        if tag == "cat32":
            continue
        # 10-step respaced loop (BASELINE config 1) with replayed noise; CFG wrapper for the img model (scale 4)
        diff = SpacedDiffusion(use_timesteps=space_timesteps(1000, [10]), betas=gd.get_named_beta_schedule("cosine", 1000, 1.),
                               model_mean_type=gd.ModelMeanType.START_X, model_var_type=gd.ModelVarType.FIXED_SMALL,
                               loss_type=gd.LossType.MSE, rescale_timesteps=False, args=args)
        B = 2
        noise = torch.randn(11, B, L, generator=g)
        draws = [noise[1 + k][:, None, :].clone() for k in range(10)]
        orig = torch.randn_like
        torch.randn_like = lambda ref, *a, **k: draws.pop(0)
        mk = {"y": {}}
        m = model
        if cond == "img":
            mk = {"y": {"context": ctx[:B], "scale": torch.ones(B) * 4.0}}
            # ClassifierFreeSampleModel asserts cond_mode == 'text' (cfg_sampler.py:21) while MDM.forward must take the
            # `context` branch (the text branch would need CLIP weights; the arithmetic is identical):
            class _Mode(str):
                def __contains__(self, item):
                    return item == "img"
            model.cond_mode = _Mode("text")
            m = ClassifierFreeSampleModel(model)
        try:
            with torch.no_grad():
                res = diff.p_sample_loop(m, (B, 1, L), noise=noise[0][:, None, :].clone(), clip_denoised=False, model_kwargs=mk,
                                         skip_timesteps=0, init_image=None, progress=False, dump_steps=None, const_noise=False)
        finally:
            torch.randn_like = orig
        out[tag + "_noise"] = noise.numpy(); out[tag + "_sample"] = res.numpy()
        print(tag, "10-step sample absmax", float(res.abs().max()))
    np.savez_compressed(os.path.join(OUT, "unet.npz"), **out)


if __name__ == "__main__":
    what = sys.argv[1:] or ["decoder"]
    for w in what:
        globals()["golden_" + w]()
