#!/usr/bin/env python
"""Decoder layer chain: one cooperative launch per pass (tc_chain_kernel) vs one launch per layer, over point-chunk sizes.
Times a 256^3 GridFiller lattice and checks the two modes agree."""
import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch
sys.path.insert(0, ".")
from surfd_b200 import synth
from surfd_b200.decoder import UdfDecoder

L, N = 32, int(sys.argv[1]) if len(sys.argv) > 1 else 256
lat = torch.randn(L, generator=torch.Generator().manual_seed(0)).cuda() * 0.7
sd = synth.synth_ae_poly(L)["decoder"]
ref = None
for chain, tiles in ((0, 6), (1, 3), (1, 6), (1, 9), (1, 12), (0, 9)):
    dec = UdfDecoder(sd, L, max_chunk_points=140 * 128 * tiles)
    dec.set_precision(1); dec.set_sm_budget(140); dec.set_chain(bool(chain))
    dec.set_latent(lat)
    best = 1e9
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        udf, grads, counts = dec.lattice(N, True)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    if ref is None:
        ref = (udf.clone(), grads.clone())
    du = float((udf - ref[0]).abs().max()); dg = float((grads - ref[1]).abs().max())
    print("chain=%d chunk=%6d points: lattice %d^3 %.1f ms  (queries %s)  max|d udf| %.2e max|d grad| %.2e" % (chain, 140 * 128 * tiles, N, best * 1e3, counts, du, dg), flush=True)
    del dec
