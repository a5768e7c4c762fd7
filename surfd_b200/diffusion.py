"""Drop-in objects for the reference's sampler boundary, so that the `sample/generate_*.py` flow of
the reference runs with only its imports swapped (see surfd_b200/compat/ for the module paths the scripts import):

  utils/model_util.py:6-16      load_model_wo_clip, create_model_and_diffusion, create_gaussian_diffusion
  models/mdm.py:11-113          MDM (constructor arguments, forward dispatch on cond_mode, eval() returning None -- F10)
  models/cfg_sampler.py:8-26    ClassifierFreeSampleModel
  diffusion/respace.py:63-132   SpacedDiffusion (timestep map, re-derived betas)
  diffusion/gaussian_diffusion.py:570-633  p_sample_loop(model, shape, noise=None, clip_denoised=..., model_kwargs=...)

These are host-side shells: every forward / sampling call goes through the C ABI (surfd_unet_forward / surfd_sample).
There is no CPU path -- constructing the device handle without the CUDA library or without a GPU raises.
"""
import collections

import torch

from . import unet as U

_Incompatible = collections.namedtuple("_IncompatibleKeys", ["missing_keys", "unexpected_keys"])


def get_named_beta_schedule(schedule_name, num_diffusion_timesteps, scale_betas=1.0):
    """diffusion/gaussian_diffusion.py:23-54"""
    if schedule_name == "linear":
        return U.linear_betas(num_diffusion_timesteps, scale_betas)
    if schedule_name == "cosine":
        return U.cosine_betas(num_diffusion_timesteps)
    raise NotImplementedError(f"unknown beta schedule: {schedule_name}")


space_timesteps = U.space_timesteps


class MDM:
    """models/mdm.py MDM: owns the denoiser weights (state_dict layout 'Unet.*') and dispatches on cond_mode."""

    def __init__(self, modeltype="", num_actions=9, dropout=0.1, activation="gelu", legacy=False, dataset="deepfasion3d",
                 clip_dim=512, arch="OpenUNet", clip_version=None, **kargs):
        if arch != "OpenUNet":
            raise NotImplementedError("only arch='OpenUNet' exists in the reference (models/mdm.py:33)")
        self.modeltype, self.num_actions, self.dataset, self.arch = modeltype, num_actions, dataset, arch
        self.dropout, self.activation, self.legacy, self.clip_dim = dropout, activation, legacy, clip_dim
        self.cond_mode = kargs.get("cond_mode", "no_cond")
        self.cond_mask_prob = kargs.get("cond_mask_prob", 0.0)
        self.clip_version = clip_version
        self.clip_model = None      # loaded on the first encode_text (surfd_b200.compat.clip); img / sketch conditioning arrives as y['context']
        self._state = None
        self._device = None
        self._samplers = {}         # latent length -> UNetSampler
        self.training = True

    # ---- nn.Module surface the scripts use ----
    def _arch_mode(self):
        # the packed architecture only distinguishes "has label_emb" (category) from the rest
        return "category" if "category" in self.cond_mode else ("no_cond" if self.cond_mode == "no_cond" else "img")

    def load_state_dict(self, state_dict, strict=True):
        exp = U.expected_keys(32, self._arch_mode(), self.num_actions)
        missing = [k for k in exp if k not in state_dict]
        unexpected = [k for k in state_dict if k not in exp]
        if "text" in self.cond_mode and not any(k.startswith("clip_model.") for k in state_dict):
            missing.append("clip_model.*")   # the reference's frozen CLIP is never in the checkpoint (load_model_wo_clip)
        if strict and (missing or unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict for MDM: missing {missing[:6]}, unexpected {unexpected[:6]}")
        for k, shp in exp.items():
            if k in state_dict and tuple(state_dict[k].shape) != shp:
                raise RuntimeError(f"size mismatch for {k}: {tuple(state_dict[k].shape)} vs {shp}")
        self._state = {k: v for k, v in state_dict.items() if k in exp}
        self._samplers.clear()
        return _Incompatible(missing, unexpected)

    def state_dict(self):
        return dict(self._state or {})

    def to(self, device):
        self._device = torch.device(device)
        if self._device.type != "cuda":
            raise RuntimeError("surfd_b200 has no CPU path: MDM.to() needs a CUDA device")
        self._samplers.clear()
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", torch.cuda.current_device() if device is None else device))

    def train(self, mode=True):
        self.training = bool(mode)   # (models/mdm.py:112-113 returns None: scripts never chain .eval())

    def eval(self):
        return self.train(False)

    def parameters(self):
        dev = self._device or torch.device("cpu")
        yield torch.empty(0, device=dev)     # `next(model.parameters()).device` (gaussian_diffusion.py:659)

    def parameters_wo_clip(self):
        return list(self.parameters())

    # ---- device handle ----
    def sampler(self, L, batch):
        if self._state is None:
            raise RuntimeError("MDM: load_state_dict() (load_model_wo_clip) must precede the first forward")
        if self._device is None:
            raise RuntimeError("MDM: call .to(device) before the first forward (surfd_b200 has no CPU path)")
        s = self._samplers.get(L)
        if s is None or s.max_batch < batch:
            if s is not None:
                s.close()
            with torch.cuda.device(self._device):
                s = U.UNetSampler(self._state, L, self._arch_mode(), self.num_actions, device=self._device, max_batch=max(8, batch))
            self._samplers[L] = s
        return s

    def encode_text(self, raw_text):
        """mdm.py:86-89: clip_model.encode_text(clip.tokenize(raw_text, truncate=True)).float() -- which the reference re-runs
        inside every one of the 1000 denoiser calls (mdm.py:96-97); here the embedding of a prompt list is computed once and
        kept.  The weights are not shipped: surfd_b200.compat.clip.load finds them through $SURFD_CLIP_PATH."""
        key = tuple(raw_text) if not isinstance(raw_text, str) else (raw_text,)
        hit = self._text_cache.get(key) if hasattr(self, "_text_cache") else None
        if hit is not None:
            return hit
        from .compat import clip
        if self.clip_model is None:
            try:
                self.clip_model, _ = clip.load(self.clip_version or "ViT-B/32", device=self._device or "cuda", jit=False)
            except RuntimeError as e:
                raise RuntimeError("cond_mode='text': %s; or pass the 512-d embeddings as y['context'] (they are constant over the "
                                   "1000 steps)" % e)
        emb = self.clip_model.encode_text(clip.tokenize(list(key), truncate=True)).float()
        self._text_cache = {key: emb}
        return emb

    def conditioning(self, y, batch):
        """(context [B,512] or None, labels [B] or None) from the reference's model_kwargs['y'] dict (mdm.py:91-110)"""
        y = y or {}
        if "sketch" in self.cond_mode or "img" in self.cond_mode:
            return y["context"], None
        if self.cond_mode == "no_cond":
            return None, None
        if "text" in self.cond_mode:
            ctx = y["context"] if "context" in y else self.encode_text(y["text"])
            return ctx, None
        return None, y["action_text"]

    def forward(self, x, timesteps, y=None):
        B, _, L = x.shape
        ctx, lab = self.conditioning(y, B)
        return self.sampler(L, B).forward(x, timesteps, ctx, lab)

    __call__ = forward


class ClassifierFreeSampleModel:
    """models/cfg_sampler.py:8-26 (sampling-time classifier-free guidance wrapper; vestigial in the reference)"""

    def __init__(self, model):
        self.model = model
        self.cond_mode = model.cond_mode
        self.clip_version = model.clip_version

    def to(self, device):
        self.model.to(device)
        return self

    def eval(self):
        self.model.eval()
        return self

    def parameters(self):
        return self.model.parameters()

    def forward(self, x, timesteps, y=None):
        assert self.model.cond_mode in ["text", "action"]
        y_uncond = dict(y)
        y_uncond["uncond"] = True
        out = self.model(x, timesteps, y)
        out_uncond = self.model(x, timesteps, y_uncond)
        return out_uncond + (y["scale"].view(-1, 1, 1).to(out.device) * (out - out_uncond))

    __call__ = forward


class SpacedDiffusion:
    """diffusion/respace.py:63-113 over gaussian_diffusion.py:123-183: x0-prediction, FIXED_SMALL variance (the only
    configuration utils/model_util.py:32-67 builds with the default sigma_small=True)."""

    def __init__(self, use_timesteps, betas, model_mean_type="START_X", model_var_type="FIXED_SMALL", loss_type=None,
                 rescale_timesteps=False, args=None):
        if str(model_mean_type).split(".")[-1] != "START_X" or str(model_var_type).split(".")[-1] != "FIXED_SMALL":
            raise NotImplementedError("surfd_b200 implements the configuration the Surf-D scripts use: START_X, FIXED_SMALL")
        if rescale_timesteps:
            raise NotImplementedError("rescale_timesteps=True is never used by the reference scripts")
        self.use_timesteps = set(use_timesteps)
        self.original_num_steps = len(betas)
        self.schedule = U.SpacedSchedule(betas, self.use_timesteps)
        self.timestep_map = self.schedule.timestep_map
        self.num_timesteps = self.schedule.num_timesteps
        self.betas = self.schedule.betas
        for name in ("alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped", "posterior_mean_coef1",
                     "posterior_mean_coef2"):
            setattr(self, name, getattr(self.schedule, name))
        self.args = args
        self.time_con = []

    def p_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None,
                      device=None, progress=False, skip_timesteps=0, init_image=None, randomize_class=False,
                      cond_fn_with_grad=False, dump_steps=None, const_noise=False):
        """Same signature as gaussian_diffusion.py:570-633; the whole reverse process is one call into the CUDA library
        (surfd_sample).  Supported: what the five scripts pass (clip_denoised=False, no denoised_fn / cond_fn, no skipped steps,
        no init_image, no dump_steps, const_noise=False); anything else raises NotImplementedError.

        Random draws: x_T = th.randn(*shape) and one th.randn_like per step from torch's default generator of `device`, in
        the reference's order (gaussian_diffusion.py:666, :507 -- drawn at t == 0 too), so a seeded run consumes the RNG
        stream exactly like the reference."""
        unsupported = []
        if clip_denoised: unsupported.append("clip_denoised=True")
        if denoised_fn is not None: unsupported.append("denoised_fn")
        if cond_fn is not None or cond_fn_with_grad: unsupported.append("cond_fn")
        if skip_timesteps: unsupported.append("skip_timesteps")
        if init_image is not None: unsupported.append("init_image")
        if randomize_class: unsupported.append("randomize_class")
        if dump_steps is not None: unsupported.append("dump_steps")
        if const_noise: unsupported.append("const_noise (crashes in the reference too: 4-D repeat on a 3-D tensor, SURVEY F10)")
        if unsupported:
            raise NotImplementedError("p_sample_loop: outside the generation path of the Surf-D scripts: " + ", ".join(unsupported))
        assert isinstance(shape, (tuple, list))
        guided = isinstance(model, ClassifierFreeSampleModel)
        base = model.model if guided else model
        if not isinstance(base, MDM):
            raise TypeError("p_sample_loop needs a surfd_b200 MDM (or its ClassifierFreeSampleModel wrapper); arbitrary Python "
                            "models cannot run in the CUDA library and there is no CPU fallback")
        if device is None:
            device = next(model.parameters()).device
        device = torch.device(device)
        B, C, L = shape
        assert C == 1
        kw = model_kwargs or {}
        y = kw.get("y") if kw else None
        if kw and not (y is None or isinstance(y, dict)):
            raise TypeError("model_kwargs['y'] must be a dict (gaussian_diffusion.py:288)")
        ctx, lab = base.conditioning(y, B)
        guidance = 1.0
        if guided:
            assert base.cond_mode in ["text", "action"]
            scale = torch.as_tensor(y["scale"], dtype=torch.float32).reshape(-1)
            if scale.numel() not in (1, B):
                raise ValueError("y['scale'] must hold one value per sample")
            if not bool((scale == scale[0]).all()):
                raise NotImplementedError("per-sample guidance scales differ; the C ABI takes one scale per call")
            guidance = float(scale[0])
            if guidance == 1.0:
                guidance = 1.0 + 2.0 ** -20   # keep the reference's two forwards; out_u == out (F2), so the result is the same
        with torch.no_grad():
            x_T = noise.to(device) if noise is not None else torch.randn(*shape, device=device)
            draws = [x_T.reshape(B, L)] + [torch.randn_like(x_T).reshape(B, L) for _ in range(self.num_timesteps)]
            noise_all = torch.stack(draws).contiguous()
        sampler = base.sampler(L, B)
        return sampler.sample(self.schedule, noise_all, ctx, lab, guidance)

    def p_sample_loop_progressive(self, *a, **k):
        raise NotImplementedError("the reverse process runs as one device-side loop; intermediate samples are not exposed")


def load_model_wo_clip(model, state_dict):
    """utils/model_util.py:6-9"""
    missing_keys, _ = model.load_state_dict(state_dict, strict=False)
    assert all([k.startswith("clip_model.") for k in missing_keys])


def get_model_args(args):
    """utils/model_util.py:19-29"""
    return {"modeltype": "", "num_actions": args.num_actions, "dropout": 0.1, "activation": "gelu", "cond_mode": args.cond_mode,
            "arch": args.arch, "clip_version": "ViT-B/32", "dataset": args.dataset}


def create_gaussian_diffusion(args):
    """utils/model_util.py:32-67: x0 prediction, 1000 steps (hard-coded), no respacing"""
    steps = 1000
    betas = get_named_beta_schedule(args.noise_schedule, steps, 1.0)
    return SpacedDiffusion(use_timesteps=space_timesteps(steps, [steps]), betas=betas, model_mean_type="START_X",
                           model_var_type="FIXED_SMALL" if args.sigma_small else "FIXED_LARGE", rescale_timesteps=False, args=args)


def create_model_and_diffusion(args):
    """utils/model_util.py:12-16"""
    return MDM(**get_model_args(args)), create_gaussian_diffusion(args)
