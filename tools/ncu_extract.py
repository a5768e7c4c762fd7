#!/usr/bin/env python
"""Target for an ncu launch list of ONE shape's extraction at the bench configuration (TF32 decoder, GridFiller, N=256):
lattice -> marching cubes -> face filter.  Also prints event-timed stage totals (not under ncu) when run plainly."""
import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch
sys.path.insert(0, ".")
from surfd_b200 import synth
from surfd_b200.decoder import UdfDecoder
from surfd_b200.meshudf import MarchingCubes, finish_mesh

L, N = 32, int(sys.argv[1]) if len(sys.argv) > 1 else 256
CHUNK = int(os.environ.get("CHUNK", str(140 * 256)))
dec = UdfDecoder(synth.synth_ae_poly(L)["decoder"], L, max_chunk_points=CHUNK)
dec.set_precision(1); dec.set_sm_budget(140)
lat = torch.randn(L, generator=torch.Generator().manual_seed(0)).cuda() * 0.7
mc = MarchingCubes()


def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e


for it in range(2 if len(sys.argv) > 2 else 1):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e0 = ev(); dec.set_latent(lat)
    udf, grads, counts = dec.lattice(N, True); udf.clamp_(min=0)
    e1 = ev()
    v, f = mc.run_raw(udf, grads)
    e2 = ev()
    verts, faces = finish_mesh(v, f, N)
    e3 = ev()
    keep = dec.face_filter(verts, faces, N)
    e4 = ev()
    fk = faces[keep.bool()]
    e5 = ev(); torch.cuda.synchronize()
    print("chunk %d " % CHUNK + "iter %d: lattice %.1f ms (udf pts %d, grad pts %d), mc %.1f ms, finish_mesh %.1f ms, face_filter %.1f ms (%d faces, %d query pts), index %.1f ms, wall %.1f ms"
          % (it, e0.elapsed_time(e1), counts[0], counts[1], e1.elapsed_time(e2), e2.elapsed_time(e3), e3.elapsed_time(e4), faces.shape[0], 9 * faces.shape[0],
             e4.elapsed_time(e5), (time.perf_counter() - t0) * 1e3))
