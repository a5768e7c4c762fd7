#!/usr/bin/env python
"""Decoder layer chain (panel-owned, all layers per panel pair) at the metric's size: 512^3 GridFiller lattice (forward +
gradient passes) and a 2M-point forward query, best of 3, over point-chunk sizes (panels of 128 rows per CTA and launch)."""
import sys, time, json
import torch
sys.path.insert(0, ".")
from surfd_b200 import synth
from surfd_b200.decoder import UdfDecoder

L, N = 32, int(sys.argv[1]) if len(sys.argv) > 1 else 512
lat = torch.randn(L, generator=torch.Generator().manual_seed(0)).cuda() * 0.7
sd = synth.synth_ae_poly(L)["decoder"]
pts = torch.rand(2_000_000, 3, device="cuda") * 2 - 1
ref = None
for panels in [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "9,8,10,4,16").split(",")]:
    dec = UdfDecoder(sd, L, max_chunk_points=140 * 128 * panels)
    dec.set_sm_budget(140); dec.set_latent(lat); dec.set_precision(1)
    best_l, best_q = 1e9, 1e9
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        udf, grads, counts = dec.lattice(N, True)
        torch.cuda.synchronize(); best_l = min(best_l, time.perf_counter() - t0)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        q = dec.query(pts)
        torch.cuda.synchronize(); best_q = min(best_q, time.perf_counter() - t0)
    if ref is None:
        ref = (udf.clone(), grads.clone(), q.clone())
    same = bool(torch.equal(udf, ref[0]) and torch.equal(grads, ref[1]) and torch.equal(q, ref[2]))
    print(json.dumps({"panels_per_cta": panels, "chunk_points": 140 * 128 * panels, "N": N, "lattice_ms": round(best_l * 1e3, 2),
                      "query_2M_ms": round(best_q * 1e3, 2), "same_as_first": same}), flush=True)
    del dec, udf, grads, q
