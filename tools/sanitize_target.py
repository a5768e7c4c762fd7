#!/usr/bin/env python
"""Small invocations of every stage of the hot path for compute-sanitizer (tools/sanitize.sh): sampler (persistent engine,
a few DDPM steps), one GridFiller lattice, marching cubes + face filter.  argv[1]: sampler | lattice | mc | all"""
import os, sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch
sys.path.insert(0, ".")
from surfd_b200 import synth, unet as U
from surfd_b200.decoder import UdfDecoder
from surfd_b200.meshudf import MarchingCubes, finish_mesh

what = sys.argv[1] if len(sys.argv) > 1 else "all"
L, N = 32, int(os.environ.get("SAN_N", "64"))
if what in ("sampler", "all"):
    B, steps = 2, int(os.environ.get("SAN_STEPS", "3"))
    net = U.UNetSampler(synth.synth_mdm(L), L, max_batch=B)
    S = U.SpacedSchedule(U.cosine_betas(), U.space_timesteps(1000, [steps]))
    noise = torch.randn(steps + 1, B, L, generator=torch.Generator().manual_seed(1)).cuda()
    out = net.sample(S, noise)
    torch.cuda.synchronize(); net.status()
    print("sampler ok", float(out.abs().max()))
if what in ("lattice", "mc", "all"):
    dec = UdfDecoder(synth.synth_ae_poly(L)["decoder"], L)
    dec.set_precision(1)
    dec.set_latent(torch.randn(L, generator=torch.Generator().manual_seed(0)).cuda() * 0.7)
    udf, grads, counts = dec.lattice(N, True); udf.clamp_(min=0)
    torch.cuda.synchronize()
    print("lattice ok", counts)
    if what in ("mc", "all"):
        mc = MarchingCubes()
        v, f = mc.run_raw(udf, grads)
        vertices, faces = finish_mesh(v, f, N)
        keep = dec.face_filter(vertices, faces, N)
        torch.cuda.synchronize()
        print("mc ok", tuple(v.shape), tuple(f.shape), int(keep.sum()))
