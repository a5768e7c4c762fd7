#!/usr/bin/env python
"""Host wall-clock + device-event breakdown of SurfDPipeline.extract() per call site (diagnostics, not a benchmark)."""
import json, os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch
sys.path.insert(0, ".")
from surfd_b200 import synth
from surfd_b200.decoder import UdfDecoder
from surfd_b200.meshudf import MarchingCubes, finish_mesh

L, N, B = 32, int(sys.argv[1]) if len(sys.argv) > 1 else 256, 8
prec = int(os.environ.get("PREC", "1"))
dec = UdfDecoder(synth.synth_ae_poly(L)["decoder"], L, max_chunk_points=(148 - 8) * 256)
dec.set_precision(prec)
dec.set_sm_budget(int(os.environ.get("BUDGET", "140")))
gen = torch.Generator().manual_seed(0)
lats = torch.randn(B, L, generator=gen).cuda() * 0.7
mcs = [MarchingCubes() for _ in range(B)]
streams = [torch.cuda.Stream() for _ in range(B)]
main = torch.cuda.current_stream()


def run(with_mc, tag):
    rows = []
    torch.cuda.synchronize(); t_all = time.perf_counter()
    fields = {}
    for k in range(B):
        t0 = time.perf_counter(); dec.set_latent(lats[k]); t1 = time.perf_counter()
        udf, grads, counts = dec.lattice(N, True); t2 = time.perf_counter()
        udf.clamp_(min=0); e1 = torch.cuda.Event(); e1.record(main); t3 = time.perf_counter()
        if with_mc:
            streams[k].wait_event(e1); mcs[k].launch(udf, grads, streams[k])
        t4 = time.perf_counter()
        fields[k] = (udf, grads)
        rows.append(dict(k=k, set_latent=t1 - t0, lattice=t2 - t1, clamp=t3 - t2, mc_launch=t4 - t3))
    torch.cuda.synchronize(); t_lat = time.perf_counter() - t_all
    t_f = time.perf_counter()
    if with_mc:
        for k in range(B):
            t0 = time.perf_counter(); res = mcs[k].finish(); t1 = time.perf_counter()
            v, f = finish_mesh(res[0], res[1], N); t2 = time.perf_counter()
            dec.set_latent(lats[k]); keep = dec.face_filter(v, f, N); t3 = time.perf_counter()
            fk = f[keep.bool()]; t4 = time.perf_counter()
            rows[k].update(mc_finish=t1 - t0, finish_mesh=t2 - t1, face_filter=t3 - t2, index=t4 - t3, faces=int(f.shape[0]))
    torch.cuda.synchronize(); t_fin = time.perf_counter() - t_f
    print(tag, "lattice phase wall %.3f s, finish phase wall %.3f s" % (t_lat, t_fin))
    for r in rows:
        print("   ", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items()})


for it in range(2):
    run(False, "no-mc  iter%d" % it)
for it in range(3):
    run(True, "with-mc iter%d" % it)

# --- does a running replay slow an unrelated GEMM? -------------------------------------------------------------
dec.set_latent(lats[0])
udf, grads, counts = dec.lattice(N, True)
udf.clamp_(min=0)
torch.cuda.synchronize()
for prec_mode in (1, 0):
    dec.set_precision(prec_mode)
    ms_idle, m = dec.time_layer(50)
    for nmc in (1, 8):
        for k in range(nmc):
            mcs[k].launch(udf, grads, streams[k])
        time.sleep(0.05)
        ms_busy, _ = dec.time_layer(50)
        for k in range(nmc):
            mcs[k].finish()
        print("precision %d: layer GEMM %.4f ms idle, %.4f ms with %d replays running" % (prec_mode, ms_idle, ms_busy, nmc))
dec.set_precision(prec)
