"""CPU suite: sampler-side host logic and the UNet oracle against the reference-generated goldens."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import unet_oracle as UO
from surfd_b200 import synth, unet as U

CASES = (("uncond32", 32, "no_cond"), ("img64", 64, "img"), ("cat32", 32, "category"))


@pytest.mark.parametrize("case", CASES, ids=lambda c: c[0])
def test_unet_oracle_matches_reference_golden(case):
    tag, L, cond = case
    g = np.load(os.path.join(GOLDEN, "unet.npz"))
    sd = synth.synth_mdm(L, cond)
    ctx = torch.from_numpy(g[tag + "_ctx"]) if cond == "img" else None
    lab = torch.from_numpy(g[tag + "_lab"]) if cond == "category" else None
    with torch.no_grad():
        o = UO.unet_forward(sd, torch.from_numpy(g[tag + "_x"]), torch.from_numpy(g[tag + "_t"]), ctx, lab)
    assert float((o - torch.from_numpy(g[tag + "_out"])).abs().max()) < 1e-5
    if tag == "cat32":
        return
    S = U.SpacedSchedule(U.cosine_betas(), U.space_timesteps(1000, [10]))
    with torch.no_grad():
        r = UO.p_sample_loop(sd, S, torch.from_numpy(g[tag + "_noise"]), ctx[:2] if ctx is not None else None, None,
                             4.0 if cond == "img" else 1.0)
    assert float((r - torch.from_numpy(g[tag + "_sample"])).abs().max()) < 1e-4


def test_schedule_tables_known_answers():
    """cosine schedule KATs (probed from the reference): beta_0 = 4.128e-5, beta_999 = 0.999,
    coef1[0] = 1, coef2[0] = 0, respacing [10] keeps steps 0,111,...,999."""
    b = U.cosine_betas()
    assert abs(b[0] - 4.128e-5) < 1e-8 and b[999] == 0.999
    S = U.SpacedSchedule(b, U.space_timesteps(1000, [1000]))
    assert S.timestep_map == list(range(1000)) and np.allclose(S.betas, b, rtol=1e-9)   # re-derived like respace.py:74-80
    assert S.posterior_mean_coef1[0] == 1.0 and S.posterior_mean_coef2[0] == 0.0
    assert S.posterior_log_variance_clipped[0] == S.posterior_log_variance_clipped[1]
    S10 = U.SpacedSchedule(b, U.space_timesteps(1000, [10]))
    assert S10.timestep_map == [0, 111, 222, 333, 444, 555, 666, 777, 888, 999]
    st = U.space_timesteps(300, [10, 15, 20])
    assert len(st) == 45 and {0, 99, 100, 199, 200, 299} <= st
    with pytest.raises(ValueError):
        U.space_timesteps(10, [20])


def test_arch_walk_and_packing():
    a = U.arch(32)
    assert len(a.keys) == 368 and a.emb_cols == 14112          # 368 tensors
    n_params = sum(int(np.prod(s)) for s in a.keys.values())
    assert n_params == 138_323_585                               # SURVEY 3.3 [probe]
    assert len(U.arch(32, "category").keys) == 369
    ops = [r[0] for r in a.prog]
    assert ops.count(U.OP_ATTN) == 16 and ops.count(U.OP_GN) == 22 * 2 + 16 + 1 and ops.count(U.OP_CONV) == 22 * 2 + 16 * 2 + 6
    sd = synth.synth_mdm(32)
    blob, prog, _ = U.pack_unet(sd, 32)
    assert blob.dtype == torch.float32 and prog.dtype == torch.int64 and prog.numel() == 16 + len(a.buffers) + len(a.prog) * U.REC
    # conv weights are repacked [tap][cout][cin]
    w = sd["Unet.input_blocks.1.0.in_layers.2.weight"]
    off = a.off["Unet.input_blocks.1.0.in_layers.2.weight"]
    assert torch.equal(blob[off:off + w.numel()].reshape(3, 224, 224), w.permute(2, 0, 1))
    bad = dict(sd); bad.pop("Unet.out.2.bias")
    with pytest.raises(RuntimeError):
        U.pack_unet(bad, 32)
    ok = dict(sd); ok["clip_model.whatever"] = torch.zeros(1)      # clip_model.* keys are ignored like load_model_wo_clip
    U.pack_unet(ok, 32)
