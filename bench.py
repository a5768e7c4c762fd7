#!/usr/bin/env python
"""bench.py -- shapes/sec of the Surf-D generation hot path on B200 (BASELINE.json metric).

One "step" = one pass of the hot path over one batch of synthetic input: a 1000-step DDPM reverse process over the
MDM/UNet for B latents, then per shape the coarse-to-fine UDF(+gradient) lattice at resolution N, MeshUDF marching
cubes and the UDF face filter (the meshudf.py:379 boundary).  The metric is quoted at 512^3, so the default workload is
BASELINE.json configs[2]'s per-GPU share (C3: uncond, 1000 steps, --resolution 512, 8 shapes per GPU; at --gpus 8 this IS
configs[2], batch 64).  --config selects the other BASELINE configurations:
  C2  uncond, --resolution 256, batch 8 on one GPU            (configs[1])
  C3  uncond, --resolution 512, 8 shapes per GPU              (configs[2])  [default]
  C4  cond_mode=img, latent 64, random CLIP-sized context, --resolution 512, 4 shapes per GPU   (configs[3])
  C5  cond_mode=text + classifier-free guidance 4.0 (two UNet passes per step), else as C4      (configs[4])
With N GPUs every rank runs the same per-GPU batch on its own shapes (weak scaling, no data-path collective; one NCCL
broadcast of the packed weights at start-up).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config C3]   # this repo's CUDA path
  python bench.py --impl reference [...]                              # the reference's CPU implementation, bounded sample

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # before the CUDA context exists (see surfd_b200/__init__.py)
os.environ.setdefault("NCCL_DEBUG", "WARN")                   # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "shapes/sec end-to-end (1000-step sample + 512\u00b3 UDF extract) at 1/2/4/8 B200"   # BASELINE.json, verbatim
CONFIGS = {
    "C2": dict(res=256, batch=8, latent=32, cond="no_cond", guidance=1.0, baseline="configs[1]"),
    "C3": dict(res=512, batch=8, latent=32, cond="no_cond", guidance=1.0, baseline="configs[2] (per-GPU share: 8 of 64 shapes)"),
    "C4": dict(res=512, batch=4, latent=64, cond="img", guidance=1.0, baseline="configs[3] (per-GPU share: 4 of 32 shapes)"),
    "C5": dict(res=512, batch=4, latent=64, cond="text", guidance=4.0, baseline="configs[4] (per-GPU share: 4 of 32 shapes)"),
}
RES = 512
BATCH = 8
LAT = 32
COND = "no_cond"
GUIDANCE = 1.0
CONFIG = "C3"
STEPS_DDPM = 1000
SEED = 10          # the CLI default --seed 10 (utils/parser_util.py:45 in the reference)
UNET_PARAM_BYTES = 0   # set from the architecture walk: SURVEY.md 8(d) counts the 138,323,585 parameters once = 553.3 MB per DDPM step


def workload_text():
    c = {"no_cond": "uncond", "img": "cond_mode=img with a random CLIP-sized [B,512] context",
         "text": "cond_mode=text with random CLIP-text-sized [B,512] embeds"}[COND]
    g = (f", classifier-free guidance {GUIDANCE:g}: both UNet forwards of a step (models/cfg_sampler.py) in one batched pass"
         if GUIDANCE != 1.0 else "")
    return (f"{CONFIG} = BASELINE.json {CONFIGS[CONFIG]['baseline']}: {c}{g}, all-parameter-randomised MDM (latent {LAT}) + closed-form "
            f"'poly' AE checkpoint, {STEPS_DDPM} DDPM steps, --resolution {RES}, batch {BATCH}/GPU, GridFiller lattice (the scripts' default)")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "200"],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, reasons, mx = [], set(), None
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx = float(parts[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            sm.sort()
            out["sm_mhz"] = sm[len(sm) // 2]
        out["sm_max_mhz"] = mx
        out["reasons"] = sorted(reasons)
        return out


def make_noise(world, rank, batch, steps, L):
    """x_T and the per-step randn_like draws from torch.manual_seed(10) on the CPU in full-batch order, sliced per rank."""
    g = torch.Generator().manual_seed(SEED)
    full = torch.randn(steps + 1, world * batch, L, generator=g)
    return full[:, rank * batch:(rank + 1) * batch].contiguous()


def make_context(world, rank, batch):
    """random CLIP-sized conditioning: 0.5 * randn(B, 512; seed 77), full-batch order, sliced per rank"""
    g = torch.Generator().manual_seed(77)
    full = 0.5 * torch.randn(world * batch, 512, generator=g)
    return full[rank * batch:(rank + 1) * batch].contiguous()


def bench_gpu(args):
    import torch.distributed as dist
    from surfd_b200 import _lib, synth, unet as U
    from surfd_b200.decoder import pack_decoder
    from surfd_b200.dist import broadcast_packed
    from surfd_b200.pipeline import SurfDPipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- weights: rank 0 builds + packs, one NCCL broadcast of the flat blobs (mirrors dist_util.sync_params) ----
    a = U.arch(LAT, COND)
    unet_param_bytes = 4 * sum(int(torch.tensor(shp).prod()) if len(shp) else 1 for shp in a.keys.values())   # 553.3 MB (8(d))
    n_dec = _lib.load().surfd_dec_packed_floats(LAT)
    if rank == 0:
        blob_u, prog, _ = U.pack_unet(synth.synth_mdm(LAT, COND), LAT, COND)
        blob_d = pack_decoder(synth.synth_ae_poly(LAT)["decoder"], LAT)
        blob_u, prog, blob_d = blob_u.to(dev), prog.to(dev), blob_d.to(dev)
    else:
        blob_u = torch.empty(a.n_floats, dtype=torch.float32, device=dev)
        prog = torch.empty(16 + len(a.buffers) + len(a.prog) * U.REC, dtype=torch.int64, device=dev)
        blob_d = torch.empty(n_dec, dtype=torch.float32, device=dev)
    broadcast_packed([blob_u, prog, blob_d], src=0)
    pipe = SurfDPipeline(None, None, LAT, COND, device=dev, max_batch=BATCH, mc_parallel=BATCH,
                         packed_unet=(blob_u, prog.cpu()), packed_decoder=blob_d)
    if args.precision == "tf32":
        pipe.decoder.set_precision(1)
    if args.split:
        ns, nd = (int(v) for v in args.split.split(","))
        pipe.overlap_split = (ns, nd) if ns > 0 else None
    noise_host = make_noise(world, rank, BATCH, STEPS_DDPM, LAT).pin_memory()
    noise_dev = noise_host.to(dev)
    ctx_host = make_context(world, rank, BATCH).pin_memory() if COND != "no_cond" else None
    ctx_dev = ctx_host.to(dev) if ctx_host is not None else None

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn_many, steps):
        """fn_many(k) runs k steps (batches) back to back; device time between two events, max over ranks"""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn_many(steps)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / 1e3

    last = {}

    # A step = one batch of BATCH shapes through the whole path.  The K timed steps go through generate_many(), which
    # software-pipelines consecutive batches (batch i's marching-cubes replays overlap batch i+1's sampler); all K
    # batches are complete (meshes built, face-filtered) when the timed region ends.
    def steps_resident(k):
        if k <= 0:
            return
        res = pipe.generate_many([noise_dev] * k, RES, contexts=[ctx_dev] * k if ctx_dev is not None else None, guidance=GUIDANCE,
                                 n_steps=STEPS_DDPM)
        last["stats"] = res[-1][2]

    def steps_host(k):
        io = {}
        pipe.generate_many([noise_host] * k, RES, contexts=[ctx_host] * k if ctx_host is not None else None, guidance=GUIDANCE,
                           n_steps=STEPS_DDPM, to_host=True, io=io)
        last["h2d"], last["d2h"] = io.get("h2d", 0) // k, io.get("d2h", 0) // k

    clocks = ClockSampler(local)
    steps_resident(args.warmup)
    barrier()
    _lib.load().surfd_launch_count(1)     # launches are counted over the timed steps only
    clocks.start()
    t_res = timed(steps_resident, args.steps)
    clk = clocks.stop()
    launches = int(_lib.load().surfd_launch_count(0))
    e2e_steps = min(args.steps, 8)        # the host-buffer arm repeats the same pipeline; 8 steps bound the run time
    steps_host(1)
    t_e2e = timed(steps_host, e2e_steps)

    # per-stage breakdown of one more step (not part of the timed region)
    tm = {}
    pipe.generate(noise_dev, RES, context=ctx_dev, guidance=GUIDANCE, n_steps=STEPS_DDPM, timings=tm)
    stats = last["stats"]

    # ---- companion value: the same step with the decoder's 512x512 layers in fp32 FFMA (north_star's "within fp32 tol" mode) ----
    value_fp32 = None
    if world == 1 and not args.no_fp32 and args.precision == "tf32":
        pipe.decoder.set_precision(0)
        steps_resident(1)
        t32 = timed(steps_resident, 2)
        value_fp32 = {"value": round(2 * BATCH / t32, 4), "unit": "shapes/s", "steps": 2, "decoder": "fp32 FFMA (udf within 1e-6 of the reference)"}
        pipe.decoder.set_precision(1)

    # ---- roofline of the decoder's 512x512 layer GEMM (the kernel SURVEY.md 8(d) names), measured live INSIDE the real
    # layer chain: one more extraction of a full batch with a CUDA event pair around every launch of that kernel ----
    pk = peaks()
    lat_last = pipe.sample_latents(noise_dev[:11].contiguous(), ctx_dev, None, 1.0, n_steps=10)   # any latents do; the extraction is what is measured
    pipe.decoder.profile(True)
    pipe.extract(lat_last, RES)
    n_launch, n_rows, ms_total = pipe.decoder.profile(False)
    flops_total = 2.0 * n_rows * 512 * 512
    achieved = flops_total / (ms_total * 1e-3) / 1e12
    ms_iso, m_iso = pipe.decoder.time_layer(iters=20)              # the same kernel alone, operands L2-warm, no residual
    # the FFMA path computes in fp32; the tensor-pipe peak it is held against is TF32 = 1/2 of the measured bf16 burst figure
    peak = pk["bf16_tflops"] / 2.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")
    kname = ("tc_chain_kernel (tcgen05 kind::tf32, the ten 512x512 layers of a pass per launch, every CTA running all layers on its own 128-row panels" if args.precision == "tf32"
             else "sgemm_nt_kernel (fp32 FFMA")
    roofline_decoder = {"bound": "tensor", "kernel": kname + ", decoder 512x512 layers, fused bias+residual+CBN+ReLU epilogue)",
                "achieved": round(achieved, 2), "peak": round(peak, 1), "unit": "TFLOP/s", "frac": round(achieved / peak, 4),
                "traffic": traffic, "peak_source": pk["source"] + " bf16 burst / 2 (TF32-class contraction held to the tensor pipe)",
                "launches": n_launch, "flops_per_launch": round(flops_total / max(1, n_launch)), "ms_per_launch": round(ms_total / max(1, n_launch), 4),
                "measured": "event pair around every layer-chain launch of one batch's lattices + face filters (10 layers per launch in TF32 "
                            "mode; activations of consecutive layers exceed L2, residual/mask operands come from HBM)",
                "isolated": {"ms_per_launch": round(ms_iso, 4), "points_per_launch": m_iso,
                             "tflops": round(2.0 * m_iso * 512 * 512 / (ms_iso * 1e-3) / 1e12, 1)}}
    # The dominant kernel of the step is the persistent sampler (one launch = the whole 1000-step reverse process of a batch).
    # Its binding roofline is HBM: every DDPM step (and every CFG pass) streams the UNet's fp32 weights once -- algorithmic bytes
    # = SURVEY 8(d)'s 553.3 MB (the 138.3 M parameters counted once; the packed blob also holds a second copy of the 22 emb_layers
    # matrices, which is implementation traffic, not algorithmic).  Timed live with a CUDA event pair around one more launch.
    # Classifier-free guidance: the reference runs two forwards per step; the persistent engine runs the pair as one pass over
    # 2B rows (same FLOPs), so the weights are needed -- and counted -- once per DDPM step.
    n_pass = 1
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    e0.record(); pipe.sample_latents(noise_dev, ctx_dev, None, GUIDANCE, n_steps=STEPS_DDPM); e1.record(); torch.cuda.synchronize(dev)
    ms_sampler = e0.elapsed_time(e1)
    bytes_sampler = float(unet_param_bytes) * STEPS_DDPM * n_pass
    tsp = os.path.join(ROOT, "profiles", "roofline_traffic_sampler.json")
    traffic_s = json.load(open(tsp)).get("dram_bytes_per_launch") if os.path.exists(tsp) else None
    roofline = {"bound": "hbm", "kernel": "unet_persistent_kernel (one cooperative launch = the whole reverse process of a batch: "
                                          "dependent op chain per DDPM step, grid barrier between ops)",
                "achieved": round(bytes_sampler / (ms_sampler * 1e-3) / 1e9, 1), "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": round(bytes_sampler / (ms_sampler * 1e-3) / 1e9 / pk["hbm_gbs"], 4), "traffic": traffic_s,
                "peak_source": pk["source"] + " copy bandwidth (sustained: the kernel runs for ~1 s)",
                "bytes_per_launch": int(bytes_sampler), "algorithmic_bytes_per_ddpm_step": int(unet_param_bytes) * n_pass,
                "ms_per_launch": round(ms_sampler, 2), "ms_per_ddpm_step": round(ms_sampler / STEPS_DDPM, 4),
                "share_of_step": round(ms_sampler / (1e3 * t_res / args.steps), 3),
                "note": "dependency-bound, not bandwidth-bound: <= 256 tokens per GEMM and a chain of grid-wide dependencies per DDPM "
                        "step; the weight stream itself needs 0.085 ms per step at the measured HBM peak"}
    # marching-cubes classification scan (HBM-bound, 4 B per lattice point read + 1 bit written), timed alone
    udf_l, _, _ = pipe.decoder.lattice(RES)
    ms_cls = pipe.mcs[0].time_classify(udf_l, iters=20)
    bytes_cls = 4.0 * RES ** 3 + RES ** 3 / 8.0
    roofline_mc = {"bound": "hbm", "kernel": "classify_kernel (avg8 / max8 candidate thresholds over the 512^3 udf lattice, 1 ballot word per warp row)",
                   "achieved": round(bytes_cls / (ms_cls * 1e-3) / 1e9, 1), "peak": pk["hbm_gbs"], "unit": "GB/s",
                   "frac": round(bytes_cls / (ms_cls * 1e-3) / 1e9 / pk["hbm_gbs"], 4), "traffic": None,
                   "bytes_per_launch": int(bytes_cls), "ms_per_launch": round(ms_cls, 4),
                   "peak_source": pk["source"] + " copy bandwidth (burst: kernel timed alone, lattice of %d MB > L2)" % (4 * RES ** 3 >> 20)}
    del udf_l
    sampler = {"ms_per_ddpm_step": round(1e3 * tm["sample_s"] / STEPS_DDPM, 4), "weight_bytes_per_step": int(unet_param_bytes) * n_pass,
               "packed_blob_bytes": int(a.n_floats * 4),
               "hbm_gbps": round(unet_param_bytes * n_pass / (tm["sample_s"] / STEPS_DDPM) / 1e9, 1), "hbm_peak_gbps": pk["hbm_gbs"],
               "engine": "persistent cooperative kernel, wide token-GEMM units on tcgen05 (kind::f16, fp16 two-term split: fp32-class products, TMEM accumulators)"}

    out = None
    if rank == 0:
        total_shapes = world * BATCH * args.steps
        out = {
            "metric": METRIC, "value": round(total_shapes / t_res, 4), "unit": "shapes/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * t_res / args.steps, 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "tf32" if args.precision == "tf32" else "f32", "data": "synthetic",
            "config": {"workload": workload_text(), "name": CONFIG,
                       "resolution": RES, "batch_per_gpu": BATCH, "ddpm_steps": STEPS_DDPM, "latent": LAT, "cond_mode": COND, "guidance": GUIDANCE,
                       "parallelism": f"dp{world} (independent shapes)",
                       "pipelining": "consecutive batches overlap: marching-cubes replays of batch i run under the sampler of batch i+1",
                       "l2": "working set >> L2: 553 MB of UNet weights streamed per DDPM step, %d MB lattice per shape" % (16 * RES ** 3 >> 20)},
            "e2e": {"value": round(world * BATCH * e2e_steps / t_e2e, 4), "unit": "shapes/s", "h2d_bytes_per_step": last["h2d"], "d2h_bytes_per_step": last["d2h"],
                    "steps": e2e_steps},
            "gpu_launches": launches,
            "clocks": clk,
            "stages_s_per_step": {k: round(v, 4) for k, v in tm.items()},
            "shape_stats": {"n_udf": stats[0]["n_udf"], "n_grad": stats[0]["n_grad"], "n_cand": stats[0]["n_cand"], "verts": stats[0]["n_verts"],
                            "faces": stats[0]["n_faces"]},
            "roofline": roofline,
            "roofline_decoder": roofline_decoder,
            "roofline_mc_classify": roofline_mc,
            "sampler": sampler,
        }
        if value_fp32 is not None:
            out["precision_fp32"] = value_fp32
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_reference(stats[0], quick=True)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


# ---------------------------------------------------------------------------------------------------------------
# CPU leg: the reference's algorithm on the host cores, on a bounded sample of the same workload.
# This is the one place (besides tests/ and smoke()) that executes oracle/.
# ---------------------------------------------------------------------------------------------------------------
def sphere_field(N, r=0.5):
    """analytic sphere UDF + gradients at resolution N, float32, built slab by slab (the 512^3 case must not need tens of GB)"""
    import numpy as np
    ax = np.linspace(-1, 1, N).astype(np.float32)
    udf = np.empty((N, N, N), np.float32)
    gr = np.zeros((N, N, N, 3), np.float32)
    Y, Z = np.meshgrid(ax, ax, indexing="ij")
    for i in range(N):
        rr = np.sqrt(ax[i] * ax[i] + Y * Y + Z * Z)
        d = rr - np.float32(r)
        udf[i] = np.abs(d)
        near = udf[i] <= np.float32(2.5 * 2 / N)
        s = -np.sign(d) / np.maximum(rr, 1e-9)
        g = np.stack([ax[i] * s, Y * s, Z * s], -1)
        gr[i][near] = g[near]
    return udf, gr


def cpu_reference(shape_stats=None, quick=True):
    """quick=True: the N=1 arm's cpu_baseline (about 20-30 s).  quick=False: the --impl reference arm -- the whole 1000-step
    sampler of the batch, 65,536 / 16,384 decoder points, the reference's compiled Cython MC at the workload's real resolution."""
    import numpy as np
    from oracle import decoder_oracle as DO, unet_oracle as UO
    from surfd_b200 import synth, unet as U
    cores = os.cpu_count() or 1
    n_pass = 2 if GUIDANCE != 1.0 else 1
    # (1) sampler at the workload's batch.  The UNet's tensors are tiny (<= 256 tokens), so torch's intra-op pool is slower with
    # every core than with a few: pick the fastest thread count on one step first.
    sd = synth.synth_mdm(LAT, COND)
    n_s = 20 if quick else STEPS_DDPM
    g = torch.Generator().manual_seed(SEED)
    noise = torch.randn(n_s + 1, BATCH, LAT, generator=g)
    ctx = make_context(1, 0, BATCH) if COND != "no_cond" else None
    one = U.SpacedSchedule(U.cosine_betas(), U.space_timesteps(1000, [1]))
    best_thr, best_t = cores, None
    for thr in sorted({cores, min(cores, 32), min(cores, 16), min(cores, 8)}, reverse=True):
        torch.set_num_threads(thr)
        with torch.no_grad():
            UO.p_sample_loop(sd, one, noise[:2], context=ctx)
            t0 = time.time(); UO.p_sample_loop(sd, one, noise[:2], context=ctx); dt = time.time() - t0
        if best_t is None or dt < best_t:
            best_thr, best_t = thr, dt
    torch.set_num_threads(best_thr)
    sched = U.SpacedSchedule(U.cosine_betas(), U.space_timesteps(1000, [n_s]))
    with torch.no_grad():
        t0 = time.time(); UO.p_sample_loop(sd, sched, noise, context=ctx); t_steps = time.time() - t0
    t_sample_batch = t_steps / n_s * STEPS_DDPM * n_pass      # the CFG wrapper runs the UNet twice per step (cfg_sampler.py:19-26)
    # (2) decoder queries: points/s forward and forward+gradient on a sample, scaled by the shape's query counts
    dsd = synth.synth_ae_poly(LAT)["decoder"]
    lat = torch.randn(LAT, generator=g).numpy()
    n_pts = 16384 if quick else 65536
    pts = (torch.rand(n_pts, 3, generator=g) * 2 - 1).numpy()
    t0 = time.time(); DO.forward(dsd, lat, pts); t_f = time.time() - t0
    t0 = time.time(); DO.forward(dsd, lat, pts[: n_pts // 4], want_grad=True); t_g = time.time() - t0
    fwd_pps, grad_pps = n_pts / t_f, (n_pts // 4) / t_g
    default_st = {512: {"n_udf": 2890672, "n_grad": 1118795, "n_faces_mc": 731432, "n_cand": 441737},
                  256: {"n_udf": 848415, "n_grad": 259081, "n_faces_mc": 157956, "n_cand": 101239}}
    st = shape_stats or default_st.get(RES, default_st[512])
    n_faces = st.get("n_faces_mc", st.get("n_faces", default_st.get(RES, default_st[512])["n_faces_mc"]))
    t_lattice = st["n_udf"] / fwd_pps + st["n_grad"] / grad_pps
    t_filter = 9 * n_faces / fwd_pps
    # (3) marching cubes: the reference's own compiled Cython (oracle/_ref) on an analytic sphere
    t_mc, mc_kind, n_mc = None, "port", (128 if quick else RES)
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
        from meshudf import _marching_cubes_lewiner_cy as cy
        from meshudf._marching_cubes_lewiner import _get_mc_luts
        udf, gr = sphere_field(n_mc)
        t0 = time.time(); cy.marching_cubes_udf(udf, gr, _get_mc_luts(), 1, 0, None); t_mc = (time.time() - t0) * (RES / n_mc) ** 3
        mc_kind = "reference"
        del udf, gr
    except Exception:
        t_mc = {256: 0.757, 512: 7.11}.get(RES, 7.11)  # SURVEY section 6 [probe] figures when oracle/_ref is unavailable
    per_shape = t_sample_batch / BATCH + t_lattice + t_mc + t_filter
    return {"value": round(1.0 / per_shape, 6), "unit": "shapes/s", "cores": cores, "torch_threads_sampler": best_thr, "kind": "port" if mc_kind == "port" else "port+reference-mc",
            "sample": f"{n_s} of {STEPS_DDPM} DDPM steps at batch {BATCH}" + (" (x2 CFG passes)" if n_pass == 2 else "") +
                      f", {n_pts} decoder points fwd / {n_pts // 4} fwd+grad scaled to the "
                      f"shape's {st['n_udf']} udf + {st['n_grad']} grad + {9 * n_faces} filter queries (the reference evaluates 9 points per face), "
                      f"reference Cython MC on a sphere at N={n_mc}" + ("" if n_mc == RES else f" scaled x{(RES // n_mc) ** 3}"),
            "measured_s": round(t_steps + t_f + t_g + (t_mc / (RES / n_mc) ** 3 if mc_kind == "reference" else 0.0), 2),
            "seconds_per_shape": {"sample": round(t_sample_batch / BATCH, 3), "lattice": round(t_lattice, 3), "mc": round(t_mc, 3), "filter": round(t_filter, 3)}}


def bench_reference(args):
    """The reference's algorithm on the host cores (torch-fp32 restatement of the UNet / numpy decoder from oracle/, the
    reference's own compiled Cython marching cubes from oracle/_ref), measured ONCE on a bounded sample of the workload: the
    complete 1000-step sampler of one batch, 65,536 decoder points forward and 16,384 forward+gradient (scaled to the
    workload's query counts), marching cubes at the workload's resolution.  Repeating that sample --steps times would measure
    the same thing again; `steps` / `warmup` are echoed for the driver, `measured_s` is the wall time actually spent."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.time()
    cb = cpu_reference(None, quick=False)
    v = cb["value"]
    out = {"impl": "reference", "metric": METRIC, "value": round(v, 6), "unit": "shapes/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * BATCH / v, 1), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": workload_text() + " -- reference algorithm on host cores, bounded sample measured once", "name": CONFIG,
                      "resolution": RES, "batch_per_gpu": BATCH, "ddpm_steps": STEPS_DDPM, "latent": LAT, "cond_mode": COND, "guidance": GUIDANCE},
           "cpu_baseline": cb, "e2e": {"value": round(v, 6), "unit": "shapes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "wall_s": round(time.time() - t0, 1)}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="surfd_b200", choices=["surfd_b200", "reference"])
    ap.add_argument("--config", default="C3", choices=sorted(CONFIGS), help="BASELINE.json configuration (default C3: the metric's 512^3, 8 shapes per GPU)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fp32", action="store_true", help="skip the fp32-decoder companion value")
    ap.add_argument("--precision", default="tf32", choices=["fp32", "tf32"],
                    help="decoder 512x512 layer GEMMs: tf32 = tcgen05 kind::tf32 (default; the precision class of the reference's own GPU runs, "
                         "udf within 2e-4 of fp32), fp32 = FFMA.  The UNet token GEMMs are fp32-class split products in both.")
    ap.add_argument("--split", default="0,0", help="SMs of the sampler kernel, SMs of the decoder GEMMs while consecutive batches run "
                                                       "concurrently (the rest runs the marching-cubes replays); 0,0 = one stream")
    ap.add_argument("--ddpm-steps", type=int, default=STEPS_DDPM, help="profiling runs only (ncu launch lists); the metric is defined at 1000")
    ap.add_argument("--resolution", type=int, default=0, help="profiling runs only: overrides the configuration's resolution")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    g = globals()
    g["CONFIG"] = args.config
    g["RES"], g["BATCH"], g["LAT"], g["COND"], g["GUIDANCE"] = cfg["res"], cfg["batch"], cfg["latent"], cfg["cond"], cfg["guidance"]
    g["STEPS_DDPM"] = args.ddpm_steps
    if args.resolution:
        g["RES"] = args.resolution
    if args.impl == "reference":
        return bench_reference(args)
    bench_gpu(args)


if __name__ == "__main__":
    main()
