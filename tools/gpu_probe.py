#!/usr/bin/env python
"""Stage timings on the GPU box (not the benchmark of record): decoder throughput, lattice, marching cubes, sampler.
Writes gpurun_out/probe.json."""
import json, os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch
sys.path.insert(0, ".")
from surfd_b200 import synth, _lib, unet as U
from surfd_b200.decoder import UdfDecoder
from surfd_b200.meshudf import MarchingCubes, get_mesh_from_udf, DecoderUdf


def timed(fn, n=3):
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 1e3)
    return best, out


def main():
    res = {"gpu": torch.cuda.get_device_name(0)}
    L = 32
    dec = UdfDecoder(synth.synth_ae_poly(L)["decoder"], L)
    gen = torch.Generator().manual_seed(0)
    lat = torch.randn(L, generator=gen)
    dec.set_latent(lat)
    M = 37888 * 8
    pts = (torch.rand(M, 3, generator=gen) * 2 - 1).cuda()
    t, _ = timed(lambda: dec.query(pts))
    res["fwd_pts_per_s"] = M / t; res["fwd_tflops"] = M * 5.308e6 / t / 1e12
    t, _ = timed(lambda: dec.query(pts, want_grad=True))
    res["fwdbwd_pts_per_s"] = M / t; res["fwdbwd_tflops"] = M * (5.308e6 + 10.617e6) / t / 1e12
    ms, m = dec.time_layer(20)
    res["layer_gemm_ms"] = ms; res["layer_gemm_tflops"] = 2 * m * 512 * 512 / ms / 1e9
    # tensor-core (TF32 tcgen05) path
    if os.environ.get("PROBE_TC", "1") == "1":
        dtc = UdfDecoder(synth.synth_ae_poly(L)["decoder"], L)
        dtc.set_precision(1); dtc.set_latent(lat)
        t, u_tc = timed(lambda: dtc.query(pts))
        res["tc_fwd_pts_per_s"] = M / t; res["tc_fwd_tflops"] = M * 5.308e6 / t / 1e12
        res["tc_vs_fp32_maxabs"] = float((u_tc - dec.query(pts)).abs().max())
        t, _ = timed(lambda: dtc.query(pts, want_grad=True))
        res["tc_fwdbwd_tflops"] = M * (5.308e6 + 10.617e6) / t / 1e12
        ms, m = dtc.time_layer(20)
        res["tc_layer_gemm_ms"] = ms; res["tc_layer_gemm_tflops"] = 2 * m * 512 * 512 / ms / 1e9
        t, (u, g, c) = timed(lambda: dtc.lattice(256, True), n=2)
        res["tc_lattice_N256_gf_s"] = t
    mc = MarchingCubes()
    sizes = [int(a) for a in sys.argv[1:]] or [128, 256]
    for N in sizes:
        t, (u, g, c) = timed(lambda: dec.lattice(N, True), n=2)
        res[f"lattice_N{N}_gf_s"] = t; res[f"lattice_N{N}_gf_counts"] = c
        t, (v, f) = timed(lambda: mc.run_raw(u.clamp(min=0), g), n=2)
        res[f"mc_N{N}_s"] = t; res[f"mc_N{N}_VF"] = [v.shape[0], f.shape[0]]; res[f"mc_N{N}_stats"] = mc.last_stats
        t, n = timed(lambda: mc.classify(u), n=3)
        res[f"classify_call_N{N}_s"] = t
        t, out = timed(lambda: get_mesh_from_udf(DecoderUdf(dec, lat), (-1, 1), 0.1, N=N, differentiable=False, max_batch=2**16, return_stats=True), n=2)
        res[f"mesh_N{N}_s"] = t; res[f"mesh_N{N}_stats"] = out[2]
    # sampler
    net = U.UNetSampler(synth.synth_mdm(L), L, max_batch=8)
    for B in (1, 8):
        x = torch.randn(B, 1, L, generator=gen); tt = torch.full((B,), 500)
        t, _ = timed(lambda: net.forward(x, tt), n=5)
        res[f"unet_forward_B{B}_ms"] = t * 1e3
        S = U.SpacedSchedule(U.cosine_betas(), U.space_timesteps(1000, [100]))
        noise = torch.randn(101, B, L, generator=gen).cuda()
        t, _ = timed(lambda: net.sample(S, noise), n=2)
        res[f"sample_100steps_B{B}_ms_per_step"] = t * 10
    res["launches"] = int(_lib.load().surfd_launch_count(0))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(os.environ.get("PROBE_OUT", "gpurun_out/probe.json"), "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
