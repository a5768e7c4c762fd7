"""Synthetic checkpoints in the reference's on-disk layouts.

The reference ships no weights and literal random init is degenerate (zero_module sites make the UNet output 0,
zero-init CBN/fc_1 make the decoder latent-independent: SURVEY F5), so benchmarks and parity tests use:
  synth_ae_rand(L, seed)   every decoder tensor randomised           -> {"decoder": state_dict}
  synth_mdm(cond_mode, seed) every UNet tensor randomised           -> flat state_dict with "Unet." keys
Both are plain torch CPU code with fixed generators, so the GPU box, this container and the oracle all see
bit-identical tensors.
"""
import math

import torch

from .decoder import expected_keys


def _fill(shape, gen, kind):
    if kind == "weight":
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        return torch.randn(shape, generator=gen) / math.sqrt(max(fan_in, 1))
    if kind == "bias":
        return 0.02 * torch.randn(shape, generator=gen)
    raise ValueError(kind)


def synth_ae_rand(latent_dim=32, seed=4321):
    """'rand' AE checkpoint: tensors with dim>=2 ~ N(0,1)/sqrt(fan_in) (fan_in = numel of one output row),
    biases ~ 0.02 N(0,1), BN running_mean ~ 0.1 N(0,1), running_var ~ U(0.5,1.5)."""
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in expected_keys(latent_dim).items():
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.tensor(0, dtype=torch.long)
        elif k.endswith("running_mean"):
            sd[k] = 0.1 * torch.randn(shp, generator=gen)
        elif k.endswith("running_var"):
            sd[k] = 0.5 + torch.rand(shp, generator=gen)
        elif k.endswith("conv_gamma.bias"):
            sd[k] = 1.0 + 0.02 * torch.randn(shp, generator=gen)
        elif len(shp) >= 2:
            sd[k] = _fill(shp, gen, "weight")
        else:
            sd[k] = _fill(shp, gen, "bias")
    return {"epoch": 0, "encoder": {}, "decoder": sd, "optimizer": {}}


# ---------------------------------------------------------------------------------------------------------
# 'poly' AE checkpoint: an exactly-constructed decoder whose field is the UDF of a 32-face convex polytope.
#
# SURVEY.md 8(d) asks for a decoder *fitted* to an analytic UDF family so that a zero-level set exists for
# any latent.  Training is not possible offline (no GPU here, ~17 h on 8 cores), so the fit is done in closed
# form.  The ReLU residual MLP computes m(x) = max_i (n_i.x - r_i(z)) by a pairwise-max tournament
# (max(a,b) = a + relu(b-a); one ConditionalResnetBlock1d per level for 32 -> 2, signed values carried as
# (+v,-v) channel pairs so nothing needs a large offset), the last block forms d = |max(A,B)| and copies it into
# K+1 channels, and the final CBN+ReLU+fc_out realises the convex piecewise-linear
#     logit(d) = 2 - 40 d + sum_j v_j relu(k_j - d)   ~=  log((0.1-d)/d)   on (0, 0.05]
# (all hinge terms positive: no cancellation in fp32/TF32), so udf(x) = 0.1*(1-sigmoid(logit)) ~= |m(x)| near the
# surface -- the unsigned distance to the polytope (exact inside and next to faces) -- and saturates smoothly
# towards 0.1 far away.  The latent moves every face: r_i(z) = r0 - tau_i(z), tau linear in z through bn_0's
# conv_beta of block 0.  Same key set / shapes as a trained checkpoint.
# ---------------------------------------------------------------------------------------------------------
POLY_K = 96


def poly_planes(n_planes=32):
    """unit normals (Fibonacci sphere), deterministic"""
    i = torch.arange(n_planes, dtype=torch.float64) + 0.5
    phi = torch.acos(1 - 2 * i / n_planes)
    theta = math.pi * (1 + 5 ** 0.5) * i
    n = torch.stack([torch.cos(theta) * torch.sin(phi), torch.sin(theta) * torch.sin(phi), torch.cos(phi)], -1)
    return n.to(torch.float32)


def _poly_u(latent_dim, seed, amp):
    gen = torch.Generator().manual_seed(seed)
    return torch.randn(32, latent_dim, generator=gen) * (amp / math.sqrt(latent_dim))


def poly_offsets(latent, seed=4321, r0=0.5, amp=0.05):
    """r_i(z) for a latent [L] (same arithmetic the checkpoint encodes; used by tests for the exact field)."""
    U = _poly_u(latent.numel(), seed, amp)
    return r0 - U @ latent.reshape(-1).to(torch.float32)


def _poly_pl():
    """knots k_0=0 < k_1 < ... < k_K = 0.05 and the hinge weights v_j of g(d) = log((0.1-d)/d) - (2 - 40 d)"""
    k = torch.logspace(math.log10(2e-5), math.log10(0.05), POLY_K, dtype=torch.float64)
    g = torch.log((0.1 - k) / k) - (2 - 40 * k)
    g[-1] = 0.0
    xs = torch.cat([torch.zeros(1, dtype=torch.float64), k])
    ys = torch.cat([torch.tensor([14.0], dtype=torch.float64), g])
    slopes = (ys[1:] - ys[:-1]) / (xs[1:] - xs[:-1])            # slope on (x_{j-1}, x_j), j = 1..K   (all < 0)
    nxt = torch.cat([slopes[1:], torch.zeros(1, dtype=torch.float64)])
    v = nxt - slopes                                             # v_j = slope_{j+1} - slope_j  (> 0 by convexity)
    return k, v


def poly_logit(d):
    """the piecewise-linear logit(d) the checkpoint encodes (float64)"""
    k, v = _poly_pl()
    d = d.to(torch.float64)
    return 2 - 40 * d + (v[None, :] * torch.relu(k[None, :] - d.reshape(-1, 1))).sum(-1).reshape(d.shape)


POLY_CLIP = 0.17


def synth_ae_poly(latent_dim=32, seed=4321, r0=0.5, amp=0.05):
    """All intermediate quantities are kept non-negative and <= 2*POLY_CLIP (plane distances are clipped to
    [-c, c] and shifted by +c before the max-tournament), so a TF32 rounding of an activation moves it by <= 6e-5:
    the field is well conditioned for the tensor-core path as well as in fp32."""
    H, c = 512, POLY_CLIP
    sd = {}
    for key, shp in expected_keys(latent_dim).items():
        if key.endswith("num_batches_tracked"):
            sd[key] = torch.tensor(0, dtype=torch.long)
        elif key.endswith("running_var"):
            sd[key] = torch.full(shp, 1.0 - 1e-5)
        elif key.endswith("conv_gamma.bias"):
            sd[key] = torch.ones(shp)
        else:
            sd[key] = torch.zeros(shp)
    n = poly_planes(32)
    U = _poly_u(latent_dim, seed, amp)
    # level 0: plane i -> channels (2i: a_i + c, 2i+1: a_i - c), a_i = n_i.x - r0 (+ tau_i(z) through bn_0's beta);
    # A'_i = relu(a_i + c) - relu(a_i - c) = clip(a_i, -c, c) + c   in [0, 2c]
    Wp, bp = sd["decoder.fc_p.weight"], sd["decoder.fc_p.bias"]
    for i in range(32):
        for j, off in ((2 * i, c), (2 * i + 1, -c)):
            Wp[j, 0:3, 0] = n[i]
            bp[j] = -r0 + off
            sd["decoder.blocks.0.bn_0.conv_beta.weight"][j, :, 0] = U[i]
    k, v = _poly_pl()
    K = k.numel()
    base = [0, 64, 80, 88]            # first channel of level 0 (pairs), levels 1..3 (single non-negative channels)
    CH_A, CH_B, CH_D = 92, 93, 94     # level 4: A' = max of the first half, B' of the second, D = B' - A'
    CH_ABS = 95                       # K+1 copies of d = |max - c| from here
    assert CH_ABS + K + 1 <= H
    for lvl in range(4):              # blocks 0..3: 32 -> 16 -> 8 -> 4 -> 2
        n_in = 32 >> lvl
        W0 = sd[f"decoder.blocks.{lvl}.fc_0.weight"]
        W1 = sd[f"decoder.blocks.{lvl}.fc_1.weight"]

        def val(row, idx, sgn):      # add sgn * X_idx (level `lvl` value) to hidden row `row`
            if lvl == 0:
                W0[row, 2 * idx, 0] += sgn; W0[row, 2 * idx + 1, 0] -= sgn
            else:
                W0[row, base[lvl] + idx, 0] += sgn
        for p in range(n_in // 2):
            h = 2 * p                 # hidden: X_a (copy, non-negative) and relu(X_b - X_a)
            val(h, 2 * p, 1.0)
            val(h + 1, 2 * p + 1, 1.0); val(h + 1, 2 * p, -1.0)
            if lvl < 3:
                W1[base[lvl + 1] + p, h, 0] = 1; W1[base[lvl + 1] + p, h + 1, 0] = 1
            else:
                dst = CH_A if p == 0 else CH_B
                W1[dst, h, 0] = 1; W1[dst, h + 1, 0] = 1
                sgn = -1.0 if p == 0 else 1.0
                W1[CH_D, h, 0] = sgn; W1[CH_D, h + 1, 0] = sgn
    # block 4: relu(A'), relu(D) -> h+ = relu(M' - c), h- = relu(c - M'), M' = A' + relu(D);  d = h+ + h-
    W0, b0, W1 = sd["decoder.blocks.4.fc_0.weight"], sd["decoder.blocks.4.fc_0.bias"], sd["decoder.blocks.4.fc_1.weight"]
    W0[0, CH_A, 0] = 1; W0[0, CH_D, 0] = 1; b0[0] = -c
    W0[1, CH_A, 0] = -1; W0[1, CH_D, 0] = -1; b0[1] = c
    for j in range(K + 1):
        W1[CH_ABS + j, 0, 0] = 1; W1[CH_ABS + j, 1, 0] = 1
    # final CBN: channel CH_ABS: relu(d) (gamma=1, beta=0) with fc_out -40; channels CH_ABS+1+j: relu(k_j - d), fc_out v_j
    gam, bet, wout = sd["decoder.bn.conv_gamma.bias"], sd["decoder.bn.conv_beta.bias"], sd["decoder.fc_out.weight"]
    wout[0, CH_ABS, 0] = -40.0
    for j in range(K):
        gam[CH_ABS + 1 + j] = -1.0
        bet[CH_ABS + 1 + j] = float(k[j])
        wout[0, CH_ABS + 1 + j, 0] = float(v[j])
    sd["decoder.fc_out.bias"][0] = 2.0
    return {"epoch": 0, "encoder": {}, "decoder": sd, "optimizer": {}}


def poly_udf(pts, latent, seed=4321, r0=0.5, amp=0.05):
    """the field the 'poly' checkpoint encodes, evaluated directly in float64:
    0.1*(1 - sigmoid(logit_PL(|clip(m, -c, c)|))), m = max_i(n_i.x - r_i(z));  ~= |m| for |m| <= 0.05"""
    n = poly_planes(32).to(torch.float64)
    r = poly_offsets(latent, seed, r0, amp).to(torch.float64)
    m = (pts.to(torch.float64) @ n.T - r).max(-1).values
    return 0.1 * (1 - torch.sigmoid(poly_logit(m.clamp(-POLY_CLIP, POLY_CLIP).abs()))), m


def synth_mdm(L=32, cond_mode="no_cond", seed=1234, num_actions=9):
    """All-parameter-randomised diffusion checkpoint in the `model{step:09d}.pt` layout (flat state_dict, 'Unet.' keys,
    no clip_model.*): tensors with dim>=2 ~ N(0,1)/sqrt(fan_in) -- including every zero_module site, which literal
    random init would leave at 0 and make the model output identically 0 -- 1-D biases ~ 0.02 N(0,1),
    GroupNorm weights 1, label_emb ~ N(0,1)/sqrt(896)."""
    from .unet import expected_keys as unet_keys
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in unet_keys(L, cond_mode, num_actions).items():
        is_gn_weight = k.endswith(".weight") and len(shp) == 1
        if is_gn_weight:
            sd[k] = torch.ones(shp)
        elif len(shp) >= 2:
            sd[k] = _fill(shp, gen, "weight")
        else:
            sd[k] = _fill(shp, gen, "bias")
    return sd


def synth_clip(seed=77, vision=(768, 12, 32, 7), text=(512, 12), embed_dim=512, context_length=77, vocab_size=49408):
    """All-parameter-randomised CLIP checkpoint in the OpenAI state-dict layout (CLIP/clip/model.py:399-436 reads the
    configuration off these shapes).  Defaults = ViT-B/32: vision (width 768, 12 layers, patch 32, 7 x 7 grid = 224 px),
    text (width 512, 12 layers), 512-d embedding, 77 tokens, 49,408-entry vocabulary.  Matrices ~ N(0,1)/sqrt(fan_in),
    biases ~ 0.02 N(0,1), LayerNorm weights 1 + 0.1 N(0,1) and biases 0.05 N(0,1), so that every parameter matters.
    Every value is fp16-representable, like the published checkpoints (stored in half precision; build_model casts the
    matrices to fp16 on load and `clip.load(device='cpu')` widens them again)."""
    gen = torch.Generator().manual_seed(seed)
    rn = lambda *shape: torch.randn(*shape, generator=gen)
    vw, vl, patch, grid = vision
    tw, tl = text
    sd = {}

    def blocks(prefix, width, layers):
        for i in range(layers):
            p = f"{prefix}.resblocks.{i}."
            sd[p + "attn.in_proj_weight"] = rn(3 * width, width) / width ** 0.5
            sd[p + "attn.in_proj_bias"] = 0.02 * rn(3 * width)
            sd[p + "attn.out_proj.weight"] = rn(width, width) / width ** 0.5
            sd[p + "attn.out_proj.bias"] = 0.02 * rn(width)
            sd[p + "ln_1.weight"] = 1 + 0.1 * rn(width)
            sd[p + "ln_1.bias"] = 0.05 * rn(width)
            sd[p + "mlp.c_fc.weight"] = rn(4 * width, width) / width ** 0.5
            sd[p + "mlp.c_fc.bias"] = 0.02 * rn(4 * width)
            sd[p + "mlp.c_proj.weight"] = rn(width, 4 * width) / (4 * width) ** 0.5
            sd[p + "mlp.c_proj.bias"] = 0.02 * rn(width)
            sd[p + "ln_2.weight"] = 1 + 0.1 * rn(width)
            sd[p + "ln_2.bias"] = 0.05 * rn(width)

    sd["visual.class_embedding"] = rn(vw) / vw ** 0.5
    sd["visual.positional_embedding"] = rn(grid * grid + 1, vw) / vw ** 0.5
    sd["visual.proj"] = rn(vw, embed_dim) / vw ** 0.5
    sd["visual.conv1.weight"] = rn(vw, 3, patch, patch) / (3 * patch * patch) ** 0.5
    sd["visual.ln_pre.weight"] = 1 + 0.1 * rn(vw)
    sd["visual.ln_pre.bias"] = 0.05 * rn(vw)
    blocks("visual.transformer", vw, vl)
    sd["visual.ln_post.weight"] = 1 + 0.1 * rn(vw)
    sd["visual.ln_post.bias"] = 0.05 * rn(vw)
    sd["token_embedding.weight"] = 0.5 * rn(vocab_size, tw)
    sd["positional_embedding"] = 0.1 * rn(context_length, tw)
    blocks("transformer", tw, tl)
    sd["ln_final.weight"] = 1 + 0.1 * rn(tw)
    sd["ln_final.bias"] = 0.05 * rn(tw)
    sd["text_projection"] = rn(tw, embed_dim) / tw ** 0.5
    sd["logit_scale"] = torch.tensor(2.6593)
    return {k: v.half().float() for k, v in sd.items()}
