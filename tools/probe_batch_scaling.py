#!/usr/bin/env python
"""Persistent sampler: ms per DDPM step as a function of the batch per launch (diagnostic)."""
import os, sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ.setdefault("SURFD_PERSIST_MAX_BATCH", "64")
import torch
sys.path.insert(0, ".")
from surfd_b200 import synth, unet as U
L = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cond = "img" if L == 64 else "no_cond"
sd = synth.synth_mdm(L, cond)
S = U.SpacedSchedule(U.cosine_betas(), U.space_timesteps(1000, [100]))
gen = torch.Generator().manual_seed(0)
for B in (4, 8, 12, 16, 24, 32, 48):
    net = U.UNetSampler(sd, L, cond, max_batch=B)
    net.set_sampler(1, 140)
    noise = torch.randn(101, B, L, generator=gen).cuda()
    ctx = torch.randn(B, 512, generator=gen).cuda() if cond == "img" else None
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); out = net.sample(S, noise, ctx); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 100)
    net.status()
    print("L=%d B=%d: %.3f ms/step, %.4f ms per sample-step, finite=%s" % (L, B, best, best / B, bool(torch.isfinite(out).all())), flush=True)
    del net
