#!/usr/bin/env python
"""Persistent sampler at batches above 8 (diagnostic): run with SURFD_PERSIST_MAX_BATCH=64, every case under its own alarm."""
import os, sys, signal, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ.setdefault("SURFD_PERSIST_MAX_BATCH", "64")
import torch
sys.path.insert(0, ".")
from surfd_b200 import synth, unet as U

L = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cond = "img" if L == 64 else "no_cond"
sd = synth.synth_mdm(L, cond)
S = U.SpacedSchedule(U.cosine_betas(), U.space_timesteps(1000, [6]))
gen = torch.Generator().manual_seed(3)
for B in (8, 9, 12, 16, 24, 40):
    net = U.UNetSampler(sd, L, cond, max_batch=B)
    noise = torch.randn(7, B, L, generator=gen)
    ctx = torch.randn(B, 512, generator=gen) if cond == "img" else None
    net.set_sampler(0)
    t0 = time.time(); ref = net.sample(S, noise, ctx); torch.cuda.synchronize()
    print("B=%d graph engine ok %.2fs" % (B, time.time() - t0), flush=True)
    for mode, n_sms in ((2, 0), (1, 0), (1, 140)):
        net.set_sampler(mode, n_sms)
        t0 = time.time()
        out = net.sample(S, noise, ctx); torch.cuda.synchronize()
        try:
            net.status(); st = "ok"
        except Exception as e:
            st = "ABORTED " + str(e)[:60]
        print("B=%d mode %d sms %d: %s maxdiff %.3e  %.2fs" % (B, mode, n_sms, st, float((out - ref).abs().max()), time.time() - t0), flush=True)
    del net
