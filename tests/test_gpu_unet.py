"""GPU parity (-m gpu): the CUDA denoiser + reverse-diffusion loop through the C-ABI against the reference goldens.
Tolerances (fp32 FFMA path): teacher-forced forward max-abs <= 1e-4 on O(1) outputs; 10-step sample <= 1e-3."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import unet_oracle as UO
from surfd_b200 import synth, unet as U

pytestmark = pytest.mark.gpu
CASES = (("uncond32", 32, "no_cond"), ("img64", 64, "img"), ("cat32", 32, "category"))


@pytest.mark.parametrize("case", CASES, ids=lambda c: c[0])
def test_forward_and_sample_match_reference_golden(case):
    tag, L, cond = case
    g = np.load(os.path.join(GOLDEN, "unet.npz"))
    net = U.UNetSampler(synth.synth_mdm(L, cond), L, cond, max_batch=8)
    ctx = torch.from_numpy(g[tag + "_ctx"]) if cond == "img" else None
    lab = torch.from_numpy(g[tag + "_lab"]) if cond == "category" else None
    o = net.forward(torch.from_numpy(g[tag + "_x"]), torch.from_numpy(g[tag + "_t"]), ctx, lab)
    err = float((o.cpu() - torch.from_numpy(g[tag + "_out"])).abs().max())
    assert err < 1e-4, err
    if tag == "cat32":
        return
    S = U.SpacedSchedule(U.cosine_betas(), U.space_timesteps(1000, [10]))
    r = net.sample(S, torch.from_numpy(g[tag + "_noise"]), ctx[:2] if ctx is not None else None, None, 4.0 if cond == "img" else 1.0)
    err = float((r.cpu() - torch.from_numpy(g[tag + "_sample"])).abs().max())
    assert err < 1e-3, err
    # the graph is cached: a second call with the same buffers replays it and gives the same bits
    r2 = net.sample(S, torch.from_numpy(g[tag + "_noise"]), ctx[:2] if ctx is not None else None, None, 4.0 if cond == "img" else 1.0)
    assert torch.equal(r, r2)


@pytest.mark.parametrize("mode,tol", [(0, 1e-4), (1, 1e-4), (2, 5e-3)], ids=["fp32-ffma", "3xtf32-mma", "tf32-mma"])
def test_token_gemm_precision_modes(mode, tol):
    """fp32 FFMA and the 3xTF32 tensor-core split both meet the fp32 tolerance; single-pass TF32 meets the 'fast' one"""
    g = np.load(os.path.join(GOLDEN, "unet.npz"))
    net = U.UNetSampler(synth.synth_mdm(64, "img"), 64, "img", max_batch=4)
    net.set_precision(mode)
    o = net.forward(torch.from_numpy(g["img64_x"]), torch.from_numpy(g["img64_t"]), torch.from_numpy(g["img64_ctx"]))
    err = float((o.cpu() - torch.from_numpy(g["img64_out"])).abs().max())
    assert err < tol, (mode, err)
    if mode == 2:
        assert err > 1e-6, "single-pass TF32 returned fp32-exact output: the tensor-core variant did not run"


def test_lanes_do_not_change_results():
    L = 32
    net = U.UNetSampler(synth.synth_mdm(L), L, max_batch=8)
    S = U.SpacedSchedule(U.cosine_betas(), U.space_timesteps(1000, [5]))
    noise = torch.randn(6, 7, L, generator=torch.Generator().manual_seed(4))
    r1 = net.sample(S, noise)
    net.set_lanes(4)
    r4 = net.sample(S, noise)
    assert torch.equal(r1, r4)


def test_batch_independence_and_ragged_batches():
    L = 32
    sd = synth.synth_mdm(L)
    net = U.UNetSampler(sd, L, max_batch=16)
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(13, 1, L, generator=gen)
    t = torch.randint(0, 1000, (13,), generator=gen)
    full = net.forward(x, t)
    for b in (1, 5):
        part = net.forward(x[:b], t[:b])
        assert torch.equal(part, full[:b])             # samples are independent and batch-size invariant
    with torch.no_grad():
        ref = UO.unet_forward(sd, x[:4], t[:4])
    assert float((full[:4].cpu() - ref).abs().max()) < 1e-4
    with pytest.raises(ValueError):
        net.forward(torch.zeros(17, 1, L), torch.zeros(17, dtype=torch.long))


def test_hundred_step_trajectory_stays_close_to_oracle():
    L = 32
    sd = synth.synth_mdm(L)
    net = U.UNetSampler(sd, L, max_batch=4)
    S = U.SpacedSchedule(U.cosine_betas(), U.space_timesteps(1000, [100]))
    gen = torch.Generator().manual_seed(10)
    noise = torch.randn(101, 2, L, generator=gen)
    r = net.sample(S, noise)
    with torch.no_grad():
        ref = UO.p_sample_loop(sd, S, noise)
    assert torch.isfinite(r).all()
    assert float((r.cpu() - ref).abs().max()) < 1e-2


@pytest.mark.parametrize("case", CASES, ids=lambda c: c[0])
def test_persistent_sampler_is_bit_identical_to_graph_replay(case):
    """surfd_sample's two engines (one cooperative persistent kernel / CUDA-graph replay of the step kernels) share the
    chunk -> warp mapping and every summation order, so they must agree bit for bit -- for every conditioning mode, with
    the CFG double pass, for ragged batches, and for any number of resident CTAs."""
    tag, L, cond = case
    gen = torch.Generator().manual_seed(11)
    net = U.UNetSampler(synth.synth_mdm(L, cond), L, cond, max_batch=8)
    S = U.SpacedSchedule(U.cosine_betas(), U.space_timesteps(1000, [6]))
    for B, guidance in ((8, 1.0), (3, 2.5 if cond == "img" else 1.0), (1, 1.0)):
        noise = torch.randn(7, B, L, generator=gen)
        ctx = torch.randn(B, 512, generator=gen) if cond == "img" else None
        lab = torch.randint(0, 9, (B,), generator=gen) if cond == "category" else None
        net.set_sampler(0)
        ref = net.sample(S, noise, ctx, lab, guidance)
        net.set_sampler(2)
        out = net.sample(S, noise, ctx, lab, guidance)
        torch.cuda.synchronize(); net.status()
        assert torch.isfinite(out).all()
        assert torch.equal(out, ref), (tag, B, float((out - ref).abs().max()))
        net.set_sampler(2, 61)          # fewer (and an odd number of) CTAs: same bits
        out2 = net.sample(S, noise, ctx, lab, guidance)
        torch.cuda.synchronize(); net.status()
        assert torch.equal(out2, ref), (tag, B, "61 CTAs", float((out2 - ref).abs().max()))
        for n_sms in (0, 140):          # default engine: one-round K split, fp32-rounding-level differences only
            net.set_sampler(1, n_sms)
            out3 = net.sample(S, noise, ctx, lab, guidance)
            torch.cuda.synchronize(); net.status()
            assert float((out3 - ref).abs().max()) < 2e-4, (tag, B, n_sms, float((out3 - ref).abs().max()))
        net.set_sampler(0)


@pytest.mark.timeout(300)
@pytest.mark.parametrize("L,cond", [(32, "no_cond"), (64, "img")], ids=["uncond32", "img64"])
def test_persistent_sampler_large_batches(L, cond):
    """More samples per call than one round of units covers (B > 8: a token GEMM has more units than resident CTAs): the
    persistent engine -- default wide units with the weight stream, and the graph-identical units -- against the graph
    engine.  (re-added; every wait in the kernel is bounded, a protocol error aborts the run and
    status() raises.)"""
    gen = torch.Generator().manual_seed(3)
    S = U.SpacedSchedule(U.cosine_betas(), U.space_timesteps(1000, [6]))
    sd = synth.synth_mdm(L, cond)
    for B in (12, 40):
        net = U.UNetSampler(sd, L, cond, max_batch=B)
        noise = torch.randn(7, B, L, generator=gen)
        ctx = torch.randn(B, 512, generator=gen) if cond == "img" else None
        net.set_sampler(0)
        ref = net.sample(S, noise, ctx)
        for mode, n_sms, tol in ((2, 0, 0.0), (1, 0, 2e-4), (1, 140, 2e-4)):
            net.set_sampler(mode, n_sms)
            out = net.sample(S, noise, ctx)
            torch.cuda.synchronize(); net.status()
            err = float((out - ref).abs().max())
            assert torch.isfinite(out).all() and err <= tol, (B, mode, n_sms, err)
        del net


@pytest.mark.parametrize("mode", [0, 2], ids=["fp32-ffma", "tf32-mma"])
def test_persistent_sampler_precision_modes(mode):
    L = 32
    net = U.UNetSampler(synth.synth_mdm(L), L, max_batch=8)
    net.set_precision(mode)
    S = U.SpacedSchedule(U.cosine_betas(), U.space_timesteps(1000, [4]))
    noise = torch.randn(5, 8, L, generator=torch.Generator().manual_seed(5))
    net.set_sampler(0); ref = net.sample(S, noise)
    net.set_sampler(2); out = net.sample(S, noise)
    torch.cuda.synchronize(); net.status()
    assert torch.equal(out, ref), float((out - ref).abs().max())


@pytest.mark.parametrize("debug,what", [("2", "in-op K-slice exchange (no deferral)"), ("16", "GroupNorm fused into the token GEMM"),
                                        ("32", "mma.sync units instead of tcgen05")],
                         ids=["no-defer", "fused-gn", "mma-sync"])
def test_persistent_sampler_variants_agree(debug, what, monkeypatch):
    """The wide-unit engine's alternative dataflows (SURFD_UNET_DEBUG, read when the op list is built) compute the same
    network: each stays within fp32-rounding distance of the graph engine, for plain, concat (output blocks) and
    attention sites, conditioning and the CFG double pass."""
    L, cond = 64, "img"
    gen = torch.Generator().manual_seed(12)
    S = U.SpacedSchedule(U.cosine_betas(), U.space_timesteps(1000, [6]))
    sd = synth.synth_mdm(L, cond)
    ref_net = U.UNetSampler(sd, L, cond, max_batch=8)
    ref_net.set_sampler(0)
    monkeypatch.setenv("SURFD_UNET_DEBUG", debug)
    net = U.UNetSampler(sd, L, cond, max_batch=8)      # its op list is built under the variant
    for B, guidance in ((8, 1.0), (3, 2.5)):
        noise = torch.randn(7, B, L, generator=gen)
        ctx = torch.randn(B, 512, generator=gen)
        ref = ref_net.sample(S, noise, ctx, None, guidance)
        out = net.sample(S, noise, ctx, None, guidance)
        torch.cuda.synchronize(); net.status()
        assert torch.isfinite(out).all()
        assert float((out - ref).abs().max()) < 2e-4, (what, B, float((out - ref).abs().max()))



def test_persistent_sampler_reports_fp16_range_violation():
    """precision mode 1 of the persistent engine splits operands into fp16 hi + lo: an activation outside the fp16 range must be
    reported by status() (SURFD_RANGE), not silently turned into inf / NaN samples."""
    from surfd_b200 import _lib
    L = 32
    sd = dict(synth.synth_mdm(L))
    sd["Unet.input_blocks.0.0.weight"] = sd["Unet.input_blocks.0.0.weight"] * 1e7     # the raw residual stream feeds the downsample conv
    net = U.UNetSampler(sd, L, max_batch=4)
    S = U.SpacedSchedule(U.cosine_betas(), U.space_timesteps(1000, [2]))
    noise = torch.randn(3, 4, L, generator=torch.Generator().manual_seed(1))
    net.sample(S, noise)
    torch.cuda.synchronize()
    with pytest.raises(_lib.SurfdError) as ei:
        net.status()
    assert ei.value.status == 6
    ok = U.UNetSampler(synth.synth_mdm(L), L, max_batch=4)                          # a healthy network stays silent
    ok.sample(S, noise); torch.cuda.synchronize(); ok.status()
