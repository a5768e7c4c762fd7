#!/usr/bin/env python
"""Locate disagreements between the FFMA and tcgen05 decoder paths (diagnostics)."""
import os, sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch
sys.path.insert(0, ".")
from surfd_b200 import synth
from surfd_b200.decoder import UdfDecoder

L = 32
gen = torch.Generator().manual_seed(11)
lat = torch.randn(L, generator=gen)
for kind in ("poly", "rand"):
    sd = (synth.synth_ae_poly(L) if kind == "poly" else synth.synth_ae_rand(L, 4321))["decoder"]
    ex = UdfDecoder(sd, L); ex.set_latent(lat)
    fa = UdfDecoder(sd, L); fa.set_precision(1); fa.set_latent(lat)
    # 1. one layer, element by element
    for M in (128, 129, 1000, 37888, 40001):
        A = torch.relu(torch.randn(M, 512, generator=torch.Generator().manual_seed(M))).cuda()
        A = (A.view(torch.int32) & -8192).view(torch.float32)
        o0, o1 = ex.debug_layer(A, 1, 0), fa.debug_layer(A, 1, 1)
        d = (o0 - o1).abs()
        bad = (d > 1e-3 * (1 + o0.abs())).nonzero()
        print(kind, "layer M=%d maxdiff %.3e bad %d" % (M, float(d.max()), bad.shape[0]),
              "rows", sorted(set((bad[:, 0]).tolist()))[:10], "cols", sorted(set((bad[:, 1]).tolist()))[:10])
    # 2. point queries of various sizes
    for M in (1, 127, 129, 1000, 40000, 80001):
        pts = (torch.rand(M, 3, generator=torch.Generator().manual_seed(M + 7)) * 2 - 1).cuda()
        a, b = ex.query(pts), fa.query(pts)
        d = (a - b).abs()
        bad = (d > 1e-3).nonzero().flatten()
        print(kind, "query M=%d maxdiff %.3e bad %d" % (M, float(d.max()), bad.numel()), bad[:12].tolist(),
              [(round(float(a[i]), 5), round(float(b[i]), 5)) for i in bad[:6].tolist()])
        if bad.numel():
            # is it the point or the position?  re-query the bad points alone
            p2 = pts[bad[:64]]
            a2, b2 = ex.query(p2), fa.query(p2)
            print("    re-query alone: maxdiff %.3e" % float((a2 - b2).abs().max()))
    # 3. lattices
    for N in (64, 128):
        u0, g0, c0 = ex.lattice(N, True)
        u1, g1, c1 = fa.lattice(N, True)
        d = (u0 - u1).abs()
        bad = (d > 1e-3).nonzero()
        print(kind, "lattice N=%d maxdiff %.3e bad %d counts %s %s" % (N, float(d.max()), bad.shape[0], c0, c1), bad[:10].tolist(),
              [(round(float(u0[tuple(i)]), 5), round(float(u1[tuple(i)]), 5)) for i in bad[:6].tolist()])
