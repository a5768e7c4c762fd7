"""Drop-in objects for the decoder side of the reference's `udf_func` closure (sample/generate_uncond.py:52-70, 96-101):

    coords_encoder = CoordsEncoder()                                   AutoEncoder/models/coordsenc.py:7-51
    decoder = CbnDecoder(coords_encoder.out_dim, latent, 512, 5)       AutoEncoder/models/cbndec.py:106-134
    decoder.load_state_dict(ckpt["decoder"], strict=True); decoder = decoder.cuda(); decoder.eval()
    def udf_func(c): return (1 - sigmoid(decoder(coords_encoder.encode(c.unsqueeze(0)), lat).squeeze(0))) * 0.1

The positional encoding and the conditional-batch-norm MLP are ONE fused device path (csrc/decoder.cu), so `encode()`
returns a light handle around the raw coordinates and `CbnDecoder.__call__` evaluates encode + decode on the GPU through
the C ABI (surfd_dec_logits).  get_mesh_from_udf() recognises closures built from these objects (see meshudf.py).
"""
import torch

from .decoder import UdfDecoder, expected_keys


class EncodedCoords:
    """What CoordsEncoder.encode returns here: the coordinates themselves ([1, M, 3]); the 63-d encoding is computed inside the
    decoder kernels (full-range sinf/cosf, csrc/decoder.cu encode_kernel)."""

    def __init__(self, coords):
        self.coords = coords

    @property
    def shape(self):
        return tuple(self.coords.shape[:-1]) + (63,)


class CoordsEncoder:
    def __init__(self, input_dims=3, include_input=True, max_freq_log2=9, num_freqs=10, log_sampling=True, periodic_fns=None):
        if (input_dims, include_input, max_freq_log2, num_freqs, log_sampling) != (3, True, 9, 10, True) or periodic_fns is not None:
            raise NotImplementedError("surfd_b200 implements the encoder configuration of the Surf-D scripts (defaults)")
        self.input_dims, self.include_input, self.max_freq_log2, self.num_freqs, self.log_sampling = 3, True, 9, 10, True
        self.out_dim = 63

    def encode(self, inputs):
        return EncodedCoords(inputs)


class CbnDecoder:
    def __init__(self, input_dim, latent_dim, hidden_dim, num_hidden_layers, out_dim=1, refine=False):
        if (input_dim, hidden_dim, num_hidden_layers, out_dim, bool(refine)) != (63, 512, 5, 1, False):
            raise NotImplementedError("surfd_b200 implements CbnDecoder(63, latent, 512, 5) (the Surf-D scripts' decoder)")
        self.latent_dim = int(latent_dim)
        self._state = None
        self._device = None
        self._dec = None
        self.training = True

    def load_state_dict(self, state_dict, strict=True):
        exp = expected_keys(self.latent_dim)
        missing = [k for k in exp if k not in state_dict]
        unexpected = [k for k in state_dict if k not in exp]
        if strict and (missing or unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict for CbnDecoder: missing {missing[:6]}, unexpected {unexpected[:6]}")
        self._state = state_dict
        self._dec = None

    def to(self, device):
        self._device = torch.device(device)
        if self._device.type != "cuda":
            raise RuntimeError("surfd_b200 has no CPU path: CbnDecoder needs a CUDA device")
        self._dec = None
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", torch.cuda.current_device() if device is None else device))

    def eval(self):
        self.training = False
        return self

    def parameters(self):
        return iter(())

    @property
    def udf_decoder(self) -> UdfDecoder:
        if self._dec is None:
            if self._state is None or self._device is None:
                raise RuntimeError("CbnDecoder: load_state_dict() and .cuda() must precede the first evaluation")
            with torch.cuda.device(self._device):
                self._dec = UdfDecoder(self._state, self.latent_dim, device=self._device)
        return self._dec

    def __call__(self, coords_emb, latent_codes):
        """coords_emb = CoordsEncoder.encode(c[None]) ; latent_codes [1, L] -> logits [1, M]"""
        if not isinstance(coords_emb, EncodedCoords):
            raise TypeError("surfd_b200 CbnDecoder takes the handle returned by surfd_b200 CoordsEncoder.encode()")
        c = coords_emb.coords
        if c.dim() != 3 or c.shape[0] != 1 or c.shape[2] != 3:
            raise ValueError("expected coordinates of shape [1, M, 3] (one shape per call, like udf_func)")
        if latent_codes.numel() != self.latent_dim:
            raise ValueError("expected one latent code of length %d" % self.latent_dim)
        dec = self.udf_decoder
        dec.set_latent(latent_codes)
        return dec.logits(c[0]).unsqueeze(0)

    forward = __call__
