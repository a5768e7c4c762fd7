#!/usr/bin/env python
"""What does a long-running one-warp kernel on another stream cost the rest of the path?  (diagnostics, not a benchmark)

Cases: the decoder layer GEMM (back-to-back launches on one stream) and the sampler's CUDA-graph step replay, each
idle / next to a spin kernel (torch.cuda._sleep) / next to marching-cubes replays; on the legacy default stream and on
a non-blocking stream."""
import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch
sys.path.insert(0, ".")
from surfd_b200 import synth, unet as U
from surfd_b200.decoder import UdfDecoder
from surfd_b200.meshudf import MarchingCubes

L, N = 32, int(sys.argv[1]) if len(sys.argv) > 1 else 256
dec = UdfDecoder(synth.synth_ae_poly(L)["decoder"], L, max_chunk_points=140 * 256)
dec.set_precision(1); dec.set_sm_budget(140)
gen = torch.Generator().manual_seed(0)
lat = torch.randn(L, generator=gen).cuda() * 0.7
dec.set_latent(lat)
udf, grads, counts = dec.lattice(N, True); udf.clamp_(min=0)
mcs = [MarchingCubes() for _ in range(8)]
side = [torch.cuda.Stream() for _ in range(8)]
work = torch.cuda.Stream()
torch.cuda.synchronize()
SPIN = int(1.9e9 * 0.4)          # ~0.4 s


def gemm_ms(stream):
    if stream is None:
        return dec.time_layer(50)[0]
    with torch.cuda.stream(stream):
        return dec.time_layer(50)[0]


def lattice_s(stream):
    ctx = torch.cuda.stream(stream) if stream is not None else torch.cuda.stream(torch.cuda.default_stream())
    with ctx:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); dec.lattice(N, True); e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / 1e3


for name, st in (("legacy", None), ("nonblocking", work)):
    print("%s stream: layer GEMM idle %.4f ms" % (name, gemm_ms(st)))
    with torch.cuda.stream(side[0]):
        torch.cuda._sleep(SPIN)
    time.sleep(0.02)
    print("%s stream: layer GEMM next to a spin kernel %.4f ms" % (name, gemm_ms(st)))
    torch.cuda.synchronize()
    mcs[0].launch(udf, grads, side[0]); time.sleep(0.02)
    print("%s stream: layer GEMM next to one replay %.4f ms" % (name, gemm_ms(st)))
    mcs[0].finish(); torch.cuda.synchronize()
    print("%s stream: lattice idle %.4f s" % (name, lattice_s(st)))
    with torch.cuda.stream(side[0]):
        torch.cuda._sleep(SPIN)
    time.sleep(0.02)
    print("%s stream: lattice next to a spin kernel %.4f s" % (name, lattice_s(st)))
    torch.cuda.synchronize()
    mcs[0].launch(udf, grads, side[0]); time.sleep(0.02)
    print("%s stream: lattice next to one replay %.4f s" % (name, lattice_s(st)))
    mcs[0].finish(); torch.cuda.synchronize()

# --- sampler graph replay next to other work ------------------------------------------------------------------------
net = U.UNetSampler(synth.synth_mdm(L), L, max_batch=8)
S = U.SpacedSchedule(U.cosine_betas(), U.space_timesteps(1000, [100]))
noise = torch.randn(101, 8, L, generator=gen).cuda()


def sample_ms(stream):
    ctx = torch.cuda.stream(stream) if stream is not None else torch.cuda.stream(torch.cuda.default_stream())
    with ctx:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); net.sample(S, noise); e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / 100


sample_ms(None); sample_ms(None)
for name, st in (("legacy", None), ("nonblocking", work)):
    print("%s stream: sampler idle %.3f ms/step" % (name, sample_ms(st)))
    with torch.cuda.stream(side[0]):
        torch.cuda._sleep(2 * SPIN)
    time.sleep(0.02)
    print("%s stream: sampler next to a spin kernel %.3f ms/step" % (name, sample_ms(st)))
    torch.cuda.synchronize()
    for k in range(8):
        mcs[k].launch(udf, grads, side[k])
    time.sleep(0.02)
    ms = sample_ms(st)
    t0 = time.perf_counter()
    for k in range(8):
        mcs[k].finish()
    torch.cuda.synchronize()
    print("%s stream: sampler next to 8 replays %.3f ms/step (replays needed %.3f s more after the sampler)" % (name, ms, time.perf_counter() - t0))

# the replays alone, and next to a running sampler
torch.cuda.synchronize(); t0 = time.perf_counter()
for k in range(8):
    mcs[k].launch(udf, grads, side[k])
for k in range(8):
    mcs[k].finish()
torch.cuda.synchronize(); print("8 replays alone: %.3f s" % (time.perf_counter() - t0))
S10 = U.SpacedSchedule(U.cosine_betas(), U.space_timesteps(1000, [1000]))
noise10 = torch.randn(1001, 8, L, generator=gen).cuda()
net.sample(S10, noise10); torch.cuda.synchronize()
t0 = time.perf_counter()
with torch.cuda.stream(work):
    net.sample(S10, noise10)
for k in range(8):
    mcs[k].launch(udf, grads, side[k])
for k in range(8):
    mcs[k].finish()
t1 = time.perf_counter() - t0
torch.cuda.synchronize(); t2 = time.perf_counter() - t0
print("8 replays next to a 1000-step sampler: replays done after %.3f s, sampler after %.3f s" % (t1, t2))
t0 = time.perf_counter()
with torch.cuda.stream(work):
    net.sample(S10, noise10)
torch.cuda.synchronize(); print("1000-step sampler alone: %.3f s" % (time.perf_counter() - t0))
