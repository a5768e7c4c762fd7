"""ORACLE (test infrastructure; never imported by the product path).

Plain PyTorch fp32 (CPU) restatement of the reference's denoiser and reverse-diffusion step, written against the
flat checkpoint tensors (no nn.Module):
  UNetModel.forward        models/openaimodel.py:710-749      ResBlock._forward   :255-275
  AttentionBlock._forward  models/openaimodel.py:318-324      QKVAttentionLegacy  :356-372
  Upsample / Downsample    models/openaimodel.py:91-119, 134-160
  timestep_embedding       utils/ldm_utils.py:165-185         GroupNorm32(32, C)  utils/ldm_utils.py:244-249
  p_sample / q_posterior   diffusion/gaussian_diffusion.py:471-520, 234-256 (x0-prediction, FIXED_SMALL, no clipping)
Pinned against the reference modules themselves by tests/golden/make_golden.py (unet_*.npz).
"""
import math

import torch
import torch.nn.functional as F

CH_MULT = (1, 2, 4, 4)
ATTN_DS = (4, 2, 1)
MC = 224


def timestep_embedding(t, dim=224, max_period=10000):
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def _gn(sd, p, x):
    return F.group_norm(x.float(), 32, sd[p + ".weight"], sd[p + ".bias"], eps=1e-5)


def _res(sd, p, x, emb):
    h = F.conv1d(F.silu(_gn(sd, p + ".in_layers.0", x)), sd[p + ".in_layers.2.weight"], sd[p + ".in_layers.2.bias"], padding=1)
    e = F.linear(F.silu(emb), sd[p + ".emb_layers.1.weight"], sd[p + ".emb_layers.1.bias"])
    h = h + e[..., None]
    h = F.conv1d(F.silu(_gn(sd, p + ".out_layers.0", h)), sd[p + ".out_layers.3.weight"], sd[p + ".out_layers.3.bias"], padding=1)
    if p + ".skip_connection.weight" in sd:
        x = F.conv1d(x, sd[p + ".skip_connection.weight"], sd[p + ".skip_connection.bias"])
    return x + h


def _attn(sd, p, x, heads=8):
    b, c, T = x.shape
    qkv = F.conv1d(_gn(sd, p + ".norm", x), sd[p + ".qkv.weight"], sd[p + ".qkv.bias"])
    ch = c // heads
    q, k, v = qkv.reshape(b * heads, ch * 3, T).split(ch, dim=1)
    scale = 1 / math.sqrt(math.sqrt(ch))
    w = torch.einsum("bct,bcs->bts", q * scale, k * scale)
    w = torch.softmax(w.float(), dim=-1)
    a = torch.einsum("bts,bcs->bct", w, v).reshape(b, -1, T)
    return x + F.conv1d(a, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])


def unet_forward(sd, x, t, context=None, y=None):
    """x [B,1,L], t [B] (original timesteps) -> [B,1,L]"""
    P = "Unet."
    emb = F.linear(timestep_embedding(t), sd[P + "time_embed.0.weight"], sd[P + "time_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd[P + "time_embed.2.weight"], sd[P + "time_embed.2.bias"])
    if y is not None:
        emb = emb + sd[P + "label_emb.weight"][y]
    if context is not None:
        emb = emb + F.linear(context, sd[P + "sketch_emb.weight"], sd[P + "sketch_emb.bias"])
    h = F.conv1d(x.float(), sd[P + "input_blocks.0.0.weight"], sd[P + "input_blocks.0.0.bias"], padding=1)
    hs = [h]
    ds, idx = 1, 1
    for level, mult in enumerate(CH_MULT):
        for _ in range(2):
            h = _res(sd, P + f"input_blocks.{idx}.0", h, emb)
            if ds in ATTN_DS:
                h = _attn(sd, P + f"input_blocks.{idx}.1", h)
            hs.append(h); idx += 1
        if level != len(CH_MULT) - 1:
            h = F.conv1d(h, sd[P + f"input_blocks.{idx}.0.op.weight"], sd[P + f"input_blocks.{idx}.0.op.bias"], stride=2, padding=1)
            hs.append(h); idx += 1; ds *= 2
    h = _res(sd, P + "middle_block.0", h, emb)
    h = _attn(sd, P + "middle_block.1", h)
    h = _res(sd, P + "middle_block.2", h, emb)
    oidx = 0
    for level, mult in list(enumerate(CH_MULT))[::-1]:
        for i in range(3):
            h = torch.cat([h, hs.pop()], dim=1)
            h = _res(sd, P + f"output_blocks.{oidx}.0", h, emb)
            sub = 1
            if ds in ATTN_DS:
                h = _attn(sd, P + f"output_blocks.{oidx}.{sub}", h); sub += 1
            if level and i == 2:
                h = F.interpolate(h, scale_factor=2, mode="nearest")
                h = F.conv1d(h, sd[P + f"output_blocks.{oidx}.{sub}.conv.weight"], sd[P + f"output_blocks.{oidx}.{sub}.conv.bias"], padding=1)
                ds //= 2
            oidx += 1
    h = F.silu(_gn(sd, P + "out.0", h))
    return F.conv1d(h, sd[P + "out.2.weight"], sd[P + "out.2.bias"], padding=1)


def p_sample_loop(sd, schedule, noise, context=None, y=None, guidance=1.0):
    """noise [n+1,B,L] (row 0 = x_T, row 1+k = randn_like of iteration k); schedule: surfd_b200.unet.SpacedSchedule-like
    object with float64 tables + timestep_map.  Returns ([B,1,L], list of per-step x0 predictions)."""
    n = schedule.num_timesteps
    x = noise[0][:, None, :].clone()
    B = x.shape[0]
    c1 = torch.from_numpy(schedule.posterior_mean_coef1)
    c2 = torch.from_numpy(schedule.posterior_mean_coef2)
    lv = torch.from_numpy(schedule.posterior_log_variance_clipped)
    tmap = torch.tensor(schedule.timestep_map)
    for k, i in enumerate(range(n)[::-1]):
        t = torch.tensor([i] * B)
        out = unet_forward(sd, x, tmap[t], context, y)
        if guidance != 1.0:
            out_u = unet_forward(sd, x, tmap[t], context, y)
            out = out_u + guidance * (out - out_u)
        mean = c1[t].float().view(-1, 1, 1) * out + c2[t].float().view(-1, 1, 1) * x
        logvar = lv[t].float().view(-1, 1, 1)
        nz = (t != 0).float().view(-1, 1, 1)
        x = mean + nz * torch.exp(0.5 * logvar) * noise[1 + k][:, None, :]
    return x
