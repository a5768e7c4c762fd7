#!/usr/bin/env python
"""Sampler engines side by side (ms per DDPM step) + the persistent kernel's per-op-type cycle split (diagnostics)."""
import os, sys, json
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch
sys.path.insert(0, ".")
from surfd_b200 import synth, unet as U

L = 32
net = U.UNetSampler(synth.synth_mdm(L), L, max_batch=8)
S = U.SpacedSchedule(U.cosine_betas(), U.space_timesteps(1000, [100]))
gen = torch.Generator().manual_seed(0)
clk = torch.cuda.get_device_properties(0).clock_rate if hasattr(torch.cuda.get_device_properties(0), "clock_rate") else 1965000


NOISE = {B: torch.randn(101, B, L, generator=gen).cuda() for B in (8, 1)}


def run(B, mode, n_sms=0, prof=False):
    noise = NOISE[B]
    net.set_sampler(mode, n_sms)
    net.profile(prof)
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); out = net.sample(S, noise); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 100)
    net.status()
    return best, out


for prec, B in ((1, 8),):
    net.set_precision(prec)
    print("---- token-GEMM precision mode %d (0 fp32 FFMA, 1 3xTF32, 2 TF32) ----" % prec)
    if os.environ.get("SURFD_UNET_DEBUG"):
        t1, r1 = run(B, 1)
        t1b, _ = run(B, 1, 100)
        print("debug %s B=%d: persistent(wide) %.3f ms/step | on 100 CTAs %.3f" % (os.environ["SURFD_UNET_DEBUG"], B, t1, t1b))
    else:
        t0, r0 = run(B, 0)
        t2, r2 = run(B, 2)
        t1, r1 = run(B, 1)
        t1b, _ = run(B, 1, 100)
        print("B=%d: graph %.3f ms/step | persistent(graph split) %.3f | persistent(wide) %.3f | wide on 100 CTAs %.3f | maxdiff %.2e / %.2e"
              % (B, t0, t2, t1, t1b, float((r2 - r0).abs().max()), float((r1 - r0).abs().max())))
    for mode in (1,):
        t, _ = run(B, mode, 0, True)
        p = net.profile()
        net.profile(False)
        ph = net.last_gemm_phases
        if ph["units"]:
            print("    first CTA token-GEMM phases (us per unit): " + ", ".join("%s %.2f" % (k, v / ph["units"] / 1965.0) for k, v in ph.items() if k not in ("units", "chunks_warp0", "wait_cycles", "w_late_pairs")),
                  "| units per step %.1f | %.2f chunk pairs per unit, %.2f us until the first pair has landed"
                  % (ph["units"] / 100, ph["chunks_warp0"] / ph["units"], ph["wait_cycles"] / ph["units"] / 1965.0), "| weight images that were late: %d of %d pairs" % (ph["w_late_pairs"], ph["chunks_warp0"]))
        gp = net.last_gn_phases
        if gp["units"]:
            print("    first CTA GroupNorm unit phases (us per unit): " + ", ".join("%s %.2f" % (k, v / gp["units"] / 1965.0) for k, v in gp.items() if k != "units"))
        print("  mode %d profiled %.3f ms/step; per op type (us per op: body / barrier, count per step)" % (mode, t))
        for cta in ("first_cta", "last_cta"):
            row = []
            for name, (body, bar, cnt) in p[cta].items():
                if cnt:
                    row.append("%s %.1f/%.1f x%d" % (name, body / cnt / 1965.0, bar / cnt / 1965.0, cnt // 100))
            tot = sum(v[0] + v[1] for v in p[cta].values()) / 100 / 1965.0
            print("   ", cta, "| ".join(row), "| total %.0f us/step" % tot)

