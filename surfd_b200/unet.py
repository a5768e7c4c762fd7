"""Host side of the denoiser: architecture walk of the reference's MDM/UNetModel, checkpoint packing, and the
op program the CUDA interpreter (csrc/unet.cu) executes once per diffusion step.

Mirrors (reference, read-only):
  models/mdm.py:34-57            fixed UNet hyper-parameters (224 ch, mult (1,2,4,4), 2 res blocks, attention at ds 1/2/4,
                                 8 heads, dims=1, context_dim=512 -> `sketch_emb`, num_classes=9 for cond_mode 'category')
  models/openaimodel.py:443-692  module construction order == state_dict key order
  models/openaimodel.py:710-749  forward: time_embed -> (+label_emb / +sketch_emb) -> input blocks -> middle -> output blocks -> out
  utils/model_util.py:6-9        load_model_wo_clip: strict=False, only clip_model.* may be missing

Activations are channels-last [B, T, C] on the device (the reference is [B, C, T]); a k=3 convolution is then three
token-shifted GEMMs over contiguous channel vectors, and the skip concat is a two-pointer read.
"""
import ctypes
import math

import torch

from . import _lib

MODEL_CH = 224
EMB_CH = 896
CH_MULT = (1, 2, 4, 4)
NUM_RES = 2
ATTN_DS = (4, 2, 1)
HEADS = 8
CONTEXT_DIM = 512

# op codes of the device program (int64 records, see csrc/unet.cu)
OP_GN, OP_CONV, OP_ATTN, OP_INCONV, OP_OUTCONV = 1, 2, 3, 4, 5
REC = 32  # int64 slots per record


class _Arch:
    """Walks the architecture once; collects (a) expected state_dict keys/shapes, (b) the weight packing plan,
    (c) the op program."""

    def __init__(self, L, num_classes=None):
        self.L = L
        self.num_classes = num_classes
        self.keys = {}       # name -> shape (reference layout)
        self.plan = []       # (name, transform) in packing order
        self.off = {}        # name -> float offset in the packed blob
        self.n_floats = 0
        self.prog = []
        self.buffers = []    # floats per batch element
        self.emb_cols = 0    # running column offset into the batched emb_layers output
        self.emb_rows = []   # (weight name, bias name, Cout)

    # ---- bookkeeping ----
    def key(self, name, shape, transform=None):
        self.keys[name] = tuple(shape)
        n = 1
        for s in shape:
            n *= s
        self.off[name] = self.n_floats
        self.plan.append((name, transform))
        self.n_floats += (n + 3) // 4 * 4     # keep every tensor 16-byte aligned
        return self.off[name]

    def buf(self, T, C):
        self.buffers.append(T * C)
        return len(self.buffers) - 1

    def rec(self, *vals):
        r = list(vals) + [0] * (REC - len(vals))
        assert len(r) == REC
        self.prog.append(r)

    # ---- modules ----
    def conv3(self, name, cin, cout):
        w = self.key(name + ".weight", (cout, cin, 3), "k3")   # packed as [3][cout][cin]
        b = self.key(name + ".bias", (cout,))
        return w, b

    def conv1(self, name, cin, cout):
        w = self.key(name + ".weight", (cout, cin, 1), "k1")   # packed as [cout][cin]
        b = self.key(name + ".bias", (cout,))
        return w, b

    def gn(self, name, c):
        return self.key(name + ".weight", (c,)), self.key(name + ".bias", (c,))

    def op_gn(self, in1, c1, in2, c2, T, out, raw, silu, gw, gb):
        self.rec(OP_GN, in1, c1, in2, c2, T, out, raw, silu, gw, gb)

    def op_conv(self, out, N, T_out, segs, bias, emb_col, residual):
        """segs: list of (in_buf, Cin, taps, stride, upsample, T_in, w_off)"""
        flat = []
        for s in segs:
            flat += list(s)
        flat += [0] * (14 - len(flat))
        self.rec(OP_CONV, out, N, T_out, len(segs), *flat, bias, emb_col, residual)

    def resblock(self, name, x_in, cin, cout, T, skip_buf=None, c_skip=0):
        """x_in (+ optional concat partner) -> new buffer.  models/openaimodel.py:255-275"""
        ctot = cin + c_skip
        g1w, g1b = self.gn(name + ".in_layers.0", ctot)
        w1, b1 = self.conv3(name + ".in_layers.2", ctot, cout)
        ew = self.key(name + ".emb_layers.1.weight", (cout, EMB_CH))
        eb = self.key(name + ".emb_layers.1.bias", (cout,))
        g2w, g2b = self.gn(name + ".out_layers.0", cout)
        w2, b2 = self.conv3(name + ".out_layers.3", cout, cout)
        has_skip_conv = ctot != cout
        if has_skip_conv:
            ws, bs = self.conv1(name + ".skip_connection", ctot, cout)
        emb_col = self.emb_cols
        self.emb_rows.append((name + ".emb_layers.1.weight", name + ".emb_layers.1.bias", cout))
        self.emb_cols += cout
        a1 = self.buf(T, ctot)
        need_raw = skip_buf is not None and has_skip_conv
        raw = self.buf(T, ctot) if need_raw else -1
        self.op_gn(x_in, cin, skip_buf if skip_buf is not None else -1, c_skip, T, a1, raw, 1, g1w, g1b)
        h1 = self.buf(T, cout)
        self.op_conv(h1, cout, T, [(a1, ctot, 3, 1, 0, T, w1)], b1, emb_col, -1)
        a2 = self.buf(T, cout)
        self.op_gn(h1, cout, -1, 0, T, a2, -1, 1, g2w, g2b)
        out = self.buf(T, cout)
        if has_skip_conv:
            src = raw if need_raw else x_in
            # out = conv3(a2) + b2 + conv1(x) + bs : the 1x1 skip is a second K segment; its bias is folded at pack time
            self.op_conv(out, cout, T, [(a2, cout, 3, 1, 0, T, w2), (src, ctot, 1, 1, 0, T, ws)], b2, -1, -1)
            self.fold_bias = getattr(self, "fold_bias", [])
            self.fold_bias.append((name + ".out_layers.3.bias", name + ".skip_connection.bias"))
        else:
            assert skip_buf is None
            self.op_conv(out, cout, T, [(a2, cout, 3, 1, 0, T, w2)], b2, -1, x_in)
        return out

    def attnblock(self, name, x_in, c, T):
        """models/openaimodel.py:318-324, 356-372"""
        gw, gb = self.gn(name + ".norm", c)
        wq, bq = self.conv1(name + ".qkv", c, 3 * c)
        wp, bp = self.conv1(name + ".proj_out", c, c)
        a = self.buf(T, c)
        self.op_gn(x_in, c, -1, 0, T, a, -1, 0, gw, gb)
        qkv = self.buf(T, 3 * c)
        self.op_conv(qkv, 3 * c, T, [(a, c, 1, 1, 0, T, wq)], bq, -1, -1)
        att = self.buf(T, c)
        self.rec(OP_ATTN, qkv, c, T, att, HEADS)
        out = self.buf(T, c)
        self.op_conv(out, c, T, [(att, c, 1, 1, 0, T, wp)], bp, -1, x_in)
        return out

    def build(self):
        L = self.L
        P = "Unet."
        self.key(P + "time_embed.0.weight", (EMB_CH, MODEL_CH)); self.key(P + "time_embed.0.bias", (EMB_CH,))
        self.key(P + "time_embed.2.weight", (EMB_CH, EMB_CH)); self.key(P + "time_embed.2.bias", (EMB_CH,))
        if self.num_classes is not None:
            self.key(P + "label_emb.weight", (self.num_classes, EMB_CH))
        self.key(P + "sketch_emb.weight", (EMB_CH, CONTEXT_DIM)); self.key(P + "sketch_emb.bias", (EMB_CH,))
        # input blocks
        w0 = self.key(P + "input_blocks.0.0.weight", (MODEL_CH, 1, 3), "k3"); b0 = self.key(P + "input_blocks.0.0.bias", (MODEL_CH,))
        h = self.buf(L, MODEL_CH)
        self.rec(OP_INCONV, h, MODEL_CH, L, w0, b0)
        hs = [(h, MODEL_CH, L)]
        ch, ds, T, idx = MODEL_CH, 1, L, 1
        for level, mult in enumerate(CH_MULT):
            for _ in range(NUM_RES):
                h = self.resblock(P + f"input_blocks.{idx}.0", h, ch, mult * MODEL_CH, T)
                ch = mult * MODEL_CH
                if ds in ATTN_DS:
                    h = self.attnblock(P + f"input_blocks.{idx}.1", h, ch, T)
                hs.append((h, ch, T)); idx += 1
            if level != len(CH_MULT) - 1:
                w, b = self.conv3(P + f"input_blocks.{idx}.0.op", ch, ch)
                out = self.buf(T // 2, ch)
                self.op_conv(out, ch, T // 2, [(h, ch, 3, 2, 0, T, w)], b, -1, -1)
                h = out; T //= 2; ds *= 2
                hs.append((h, ch, T)); idx += 1
        # middle
        h = self.resblock(P + "middle_block.0", h, ch, ch, T)
        h = self.attnblock(P + "middle_block.1", h, ch, T)
        h = self.resblock(P + "middle_block.2", h, ch, ch, T)
        # output blocks
        oidx = 0
        for level, mult in list(enumerate(CH_MULT))[::-1]:
            for i in range(NUM_RES + 1):
                sb, sc, sT = hs.pop()
                assert sT == T
                h = self.resblock(P + f"output_blocks.{oidx}.0", h, ch, MODEL_CH * mult, T, skip_buf=sb, c_skip=sc)
                ch = MODEL_CH * mult
                sub = 1
                if ds in ATTN_DS:
                    h = self.attnblock(P + f"output_blocks.{oidx}.{sub}", h, ch, T); sub += 1
                if level and i == NUM_RES:
                    w, b = self.conv3(P + f"output_blocks.{oidx}.{sub}.conv", ch, ch)
                    out = self.buf(T * 2, ch)
                    self.op_conv(out, ch, T * 2, [(h, ch, 3, 1, 1, T, w)], b, -1, -1)
                    h = out; T *= 2; ds //= 2
                oidx += 1
        gw, gb = self.gn(P + "out.0", ch)
        wo = self.key(P + "out.2.weight", (1, MODEL_CH, 3), "k3"); bo = self.key(P + "out.2.bias", (1,))
        a = self.buf(T, ch)
        self.op_gn(h, ch, -1, 0, T, a, -1, 1, gw, gb)
        self.rec(OP_OUTCONV, a, ch, T, wo, bo)
        # batched emb_layers: one [sum Cout, 896] matrix + bias, appended after the per-tensor region
        self.emb_w_off = self.n_floats
        self.n_floats += self.emb_cols * EMB_CH
        self.emb_b_off = self.n_floats
        self.n_floats += (self.emb_cols + 3) // 4 * 4
        return self


def arch(L, cond_mode="no_cond", num_actions=9):
    num_classes = num_actions if "category" in cond_mode else None
    return _Arch(L, num_classes).build()


def expected_keys(L=32, cond_mode="no_cond", num_actions=9):
    return dict(arch(L, cond_mode, num_actions).keys)


def pack_unet(state_dict, L=32, cond_mode="no_cond", num_actions=9):
    """state_dict (flat, 'Unet.' prefix, as model{step:09d}.pt) -> (float32 blob, int64 program, arch)."""
    a = arch(L, cond_mode, num_actions)
    missing = [k for k in a.keys if k not in state_dict]
    if missing:   # load_model_wo_clip: only clip_model.* keys may be missing (utils/model_util.py:6-9)
        raise RuntimeError(f"Error(s) in loading state_dict for MDM: missing keys {missing[:8]}{'...' if len(missing) > 8 else ''}")
    # unexpected keys are ignored like the reference's load_state_dict(strict=False), which discards that list (e.g. a
    # category checkpoint's Unet.label_emb.weight sampled with --cond_mode no_cond, clip_model.*, EMA leftovers)
    blob = torch.zeros(a.n_floats, dtype=torch.float32)
    fold = dict(getattr(a, "fold_bias", []))
    for name, tr in a.plan:
        t = state_dict[name].detach().to(torch.float32).cpu()
        if tuple(t.shape) != a.keys[name]:
            raise RuntimeError(f"size mismatch for {name}: {tuple(t.shape)} vs {a.keys[name]}")
        if tr == "k3":
            t = t.permute(2, 0, 1).contiguous()       # [cout][cin][3] -> [3][cout][cin]
        elif tr == "k1":
            t = t[:, :, 0].contiguous()
        if name in fold:                               # conv bias + 1x1 skip-conv bias share one epilogue add
            t = t + state_dict[fold[name]].detach().to(torch.float32).cpu()
        blob[a.off[name]:a.off[name] + t.numel()] = t.reshape(-1)
    col = 0
    for wname, bname, cout in a.emb_rows:
        blob[a.emb_w_off + col * EMB_CH: a.emb_w_off + (col + cout) * EMB_CH] = state_dict[wname].detach().float().cpu().reshape(-1)
        blob[a.emb_b_off + col: a.emb_b_off + col + cout] = state_dict[bname].detach().float().cpu()
        col += cout
    hdr = [len(a.buffers), len(a.prog), a.emb_cols, a.emb_w_off, a.emb_b_off,
           a.off["Unet.time_embed.0.weight"], a.off["Unet.time_embed.0.bias"],
           a.off["Unet.time_embed.2.weight"], a.off["Unet.time_embed.2.bias"],
           a.off["Unet.sketch_emb.weight"], a.off["Unet.sketch_emb.bias"],
           a.off.get("Unet.label_emb.weight", -1), a.num_classes or 0, a.L, 0, 0]
    prog = torch.tensor(hdr + a.buffers + [v for r in a.prog for v in r], dtype=torch.int64)
    return blob, prog, a


def cosine_betas(n=1000, max_beta=0.999):
    """diffusion/gaussian_diffusion.py:23-67 (cosine schedule), float64"""
    import numpy as np
    f = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
    return np.array([min(1 - f((i + 1) / n) / f(i / n), max_beta) for i in range(n)], dtype=np.float64)


def linear_betas(n=1000, scale_betas=1.0):
    import numpy as np
    scale = scale_betas * 1000 / n
    return np.linspace(scale * 0.0001, scale * 0.02, n, dtype=np.float64)


def space_timesteps(num_timesteps, section_counts):
    """diffusion/respace.py:7-60"""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            desired = int(section_counts[len("ddim"):])
            for i in range(1, num_timesteps):
                if len(range(0, num_timesteps, i)) == desired:
                    return set(range(0, num_timesteps, i))
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(x) for x in section_counts.split(",")]
    size_per = num_timesteps // len(section_counts)
    extra = num_timesteps % len(section_counts)
    start, all_steps = 0, []
    for i, count in enumerate(section_counts):
        size = size_per + (1 if i < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        stride = 1 if count <= 1 else (size - 1) / (count - 1)
        cur = 0.0
        for _ in range(count):
            all_steps.append(start + round(cur))
            cur += stride
        start += size
    return set(all_steps)


class SpacedSchedule:
    """Coefficient tables of GaussianDiffusion.__init__ (gaussian_diffusion.py:123-183) for the retained steps of
    SpacedDiffusion (respace.py:63-86): float64 tables, x0-prediction, FIXED_SMALL variance."""

    def __init__(self, betas, use_timesteps):
        import numpy as np
        base_ac = np.cumprod(1.0 - np.asarray(betas, dtype=np.float64), axis=0)
        last, new_betas, tmap = 1.0, [], []
        for i, ac in enumerate(base_ac):
            if i in use_timesteps:
                new_betas.append(1 - ac / last)
                last = ac
                tmap.append(i)
        b = np.array(new_betas, dtype=np.float64)
        self.timestep_map = tmap
        self.num_timesteps = len(b)
        self.betas = b
        alphas = 1.0 - b
        ac = np.cumprod(alphas, axis=0)
        ac_prev = np.append(1.0, ac[:-1])
        self.alphas_cumprod = ac
        self.posterior_variance = b * (1.0 - ac_prev) / (1.0 - ac)
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:])) \
            if len(b) > 1 else np.log(self.posterior_variance)
        self.posterior_mean_coef1 = b * np.sqrt(ac_prev) / (1.0 - ac)
        self.posterior_mean_coef2 = (1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac)

    def device_tables(self, device):
        """[3][n] float32: coef1, coef2, exp(0.5*logvar) -- each rounded the way _extract_into_tensor(...).float() and the
        fp32 `th.exp(0.5 * log_variance)` of p_sample do (gaussian_diffusion.py:1329-1342, :519).  Cached per device so the
        sampler's captured CUDA graph (keyed on these pointers) is reused across calls."""
        key = str(device)
        cache = self.__dict__.setdefault("_dev_tables", {})
        if key in cache:
            return cache[key]
        cache[key] = self._device_tables(device)
        return cache[key]

    def _device_tables(self, device):
        c1 = torch.from_numpy(self.posterior_mean_coef1).float()
        c2 = torch.from_numpy(self.posterior_mean_coef2).float()
        lv = torch.from_numpy(self.posterior_log_variance_clipped).float()
        std = torch.exp(0.5 * lv)
        return torch.stack([c1, c2, std]).contiguous().to(device), torch.tensor(self.timestep_map, dtype=torch.int64, device=device)


class UNetSampler:
    """Device-resident denoiser + reverse-diffusion loop (one per GPU process)."""

    def __init__(self, state_dict, L=32, cond_mode="no_cond", num_actions=9, device="cuda", max_batch=64, packed=None):
        """`packed` = (blob, prog) from pack_unet() (e.g. received by NCCL broadcast) instead of a state_dict."""
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.L, self.cond_mode = L, cond_mode
        if packed is not None:
            blob, prog = packed
            self.arch = arch(L, cond_mode, num_actions)
            assert blob.numel() == self.arch.n_floats
        else:
            blob, prog, self.arch = pack_unet(state_dict, L, cond_mode, num_actions)
        self.n_weight_floats = blob.numel()
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.surfd_unet_create(_lib.ptr(blob), blob.numel(), _lib.ptr(prog), prog.numel(), L, int(max_batch),
                                                  ctypes.byref(h)))
        self._h = h
        self.max_batch = max_batch

    def close(self):
        if getattr(self, "_h", None):
            self.lib.surfd_unet_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_precision(self, mode):
        """token GEMMs: 0 fp32 FFMA; 1 (default) fp32-class split products -- 3xTF32 on mma.sync in the per-op kernels, an fp16
        two-term split (hi + lo/4096, operands must lie in the fp16 range |x| < 65504) on tcgen05 in the persistent engine,
        both 2^-22 relative; 2 single-pass TF32"""
        _lib.check(self.lib.surfd_unet_set_precision(self._h, int(mode)))

    def set_lanes(self, n):
        """number of concurrent sampler streams a batch is split over (default min(8, max_batch))"""
        _lib.check(self.lib.surfd_unet_set_lanes(self._h, int(n)))

    def set_sampler(self, mode, n_sms=0):
        """sample() engine: 1 = persistent cooperative kernel with wide token-GEMM units (default; `n_sms` CTAs, 0 = one per
        SM), 0 = CUDA-graph replay of the step kernels, 2 = persistent kernel with the graph path's units and K split
        (bit-identical to mode 0)"""
        _lib.check(self.lib.surfd_unet_set_sampler(self._h, int(mode), int(n_sms)))

    def profile(self, on=None):
        """on=True/False: switch the persistent kernel's cycle counters; on=None: read the last run's counters as
        {cta: {op: (body_cycles, barrier_cycles, count)}} -- diagnostics"""
        if on is not None:
            _lib.check(self.lib.surfd_unet_profile(self._h, int(bool(on)), None))
            return None
        buf = (ctypes.c_int64 * 64)()
        _lib.check(self.lib.surfd_unet_profile(self._h, 0, buf))
        self.last_gemm_phases = dict(chunk_stream=buf[0], cta_partial=buf[1], publish=buf[2], reduce_epilogue=buf[24], units=buf[25], chunks_warp0=buf[26], wait_cycles=buf[27],
                                     w_full_wait=buf[48], stage_wait=buf[49], done_wait=buf[50], w_late_pairs=buf[51],
                                     split=buf[52], fence=buf[53], sync=buf[54], mma_issue=buf[55], a_issue=buf[56], a_wait=buf[57])
        self.last_gn_phases = dict(load_sum=buf[58], mean_reduce=buf[59], var_reduce=buf[60], emit=buf[61], units=buf[62])
        names = {1: "emb1", 2: "linear", 3: "inconv", 4: "groupnorm", 5: "token_gemm", 6: "attention", 7: "outconv+update"}
        return {("first_cta", "last_cta")[h]: {names[o]: tuple(buf[(h * 8 + o) * 3 + i] for i in range(3)) for o in names} for h in range(2)}

    def status(self):
        """raises if the last persistent sample() aborted (call after synchronising its stream)"""
        _lib.check(self.lib.surfd_unet_status(self._h))

    def _check_cond(self, B, context, labels):
        """Host-side validation of the conditioning before anything is launched: the device kernels index
        `label_emb[y[b]]` and read B context rows unchecked.  Same failures as the reference: nn.Embedding raises IndexError
        for a label outside [0, num_classes) (openaimodel.py:727-730), a category model asserts that labels are given
        (`assert (y is not None) == (self.num_classes is not None)`, openaimodel.py:719-721), F.linear raises on a context of
        the wrong width."""
        nc = self.arch.num_classes
        if (labels is not None) != (nc is not None):
            raise ValueError("must specify y (labels) if and only if the model is class-conditional (cond_mode 'category')")
        if labels is not None:
            if labels.dim() != 1 or labels.shape[0] != B:
                raise ValueError(f"labels must have shape ({B},), got {tuple(labels.shape)}")
            if labels.is_floating_point():
                raise TypeError("labels must be an integer tensor (nn.Embedding indices)")
            lo, hi = int(labels.min()), int(labels.max())
            if lo < 0 or hi >= nc:
                raise IndexError(f"index out of range in self: label {lo if lo < 0 else hi} outside [0, {nc})")
        if context is not None:
            if context.dim() != 2 or tuple(context.shape) != (B, CONTEXT_DIM):
                raise ValueError(f"context must have shape ({B}, {CONTEXT_DIM}), got {tuple(context.shape)}")

    def forward(self, x, t, context=None, labels=None):
        """x [B,1,L], t [B] int64 (original-process timesteps) -> model output [B,1,L]  (MDM.forward)"""
        B = x.shape[0]
        self._check_cond(B, context, labels)
        x = x.detach().to(self.device, torch.float32).reshape(B, self.L).contiguous()
        t = t.detach().to(self.device, torch.int64).contiguous()
        ctx = context.detach().to(self.device, torch.float32).contiguous() if context is not None else None
        lab = labels.detach().to(self.device, torch.int64).contiguous() if labels is not None else None
        out = torch.empty(B, self.L, device=self.device, dtype=torch.float32)
        _lib.check(self.lib.surfd_unet_forward(self._h, B, _lib.ptr(x), _lib.ptr(t), _lib.ptr(ctx), _lib.ptr(lab), _lib.ptr(out),
                                               _lib.stream_ptr()))
        return out.reshape(B, 1, self.L)

    def sample(self, schedule, noise, context=None, labels=None, guidance=1.0):
        """noise [n_steps+1, B, L]: row 0 = x_T, row 1+k = randn_like of loop iteration k.  Returns [B,1,L]."""
        n = schedule.num_timesteps
        B = noise.shape[1]
        assert noise.shape[0] == n + 1 and noise.shape[2] == self.L
        self._check_cond(B, context, labels)
        coef, tmap = schedule.device_tables(self.device)
        noise = noise.detach().to(self.device, torch.float32).contiguous()
        ctx = context.detach().to(self.device, torch.float32).contiguous() if context is not None else None
        lab = labels.detach().to(self.device, torch.int64).contiguous() if labels is not None else None
        out = torch.empty(B, self.L, device=self.device, dtype=torch.float32)
        _lib.check(self.lib.surfd_sample(self._h, B, n, _lib.ptr(tmap), _lib.ptr(coef), _lib.ptr(noise), _lib.ptr(ctx), _lib.ptr(lab),
                                         float(guidance), _lib.ptr(out), _lib.stream_ptr()))
        return out.reshape(B, 1, self.L)
