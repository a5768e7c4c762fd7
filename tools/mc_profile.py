#!/usr/bin/env python
"""Marching-cubes replay timing (+ cycle breakdown when the library was built with SURFD_MC_FLAGS=-DMC_PROFILE)."""
import os, sys, time, json
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch
sys.path.insert(0, ".")
from surfd_b200 import synth
from surfd_b200.decoder import UdfDecoder
from surfd_b200.meshudf import MarchingCubes

L, N = 32, int(sys.argv[1]) if len(sys.argv) > 1 else 256
dec = UdfDecoder(synth.synth_ae_poly(L)["decoder"], L)
dec.set_precision(1)
lat = torch.randn(L, generator=torch.Generator().manual_seed(0)).cuda() * 0.7
dec.set_latent(lat)
udf, grads, counts = dec.lattice(N, True); udf.clamp_(min=0)
mc = MarchingCubes()
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    v, f = mc.run_raw(udf, grads)
    torch.cuda.synchronize(); t = time.perf_counter() - t0
    prof = mc.profile()
    out = dict(N=N, seconds=round(t, 4), V=int(v.shape[0]), F=int(f.shape[0]), stats=mc.last_stats)
    if prof["total"]:
        tot = prof["total"]
        out["cycles"] = prof
        out["share"] = {k: round(prof[k] / tot, 3) for k in ("fetch", "sign", "tiling", "emit")}
        out["share"]["queue+loop"] = round(1 - sum(out["share"].values()), 3)
        out["cycles_per_visit"] = round(tot / max(1, prof["visits"]))
    print(json.dumps(out))
