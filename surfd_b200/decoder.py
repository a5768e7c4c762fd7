"""Host side of the UDF decoder: checkpoint packing and the handle wrapper.

Mirrors the reference objects that make up `udf_func` (sample/generate_uncond.py:52-70, 96-101):
`CoordsEncoder()` + `CbnDecoder(63, latent, 512, 5)` loaded from `ckpt["decoder"]` with strict=True.
"""
import ctypes

import torch

from . import _lib

HID, ENC, NBLK = 512, 64, 5


def _cbn_prefixes():
    names = []
    for i in range(NBLK):
        names += [f"decoder.blocks.{i}.bn_0", f"decoder.blocks.{i}.bn_1"]
    names.append("decoder.bn")
    return names


def expected_keys(latent_dim):
    """The exact key set of a CbnDecoder state_dict (AutoEncoder/models/cbndec.py)."""
    keys = {"decoder.fc_p.weight": (HID, 63, 1), "decoder.fc_p.bias": (HID,),
            "decoder.fc_out.weight": (1, HID, 1), "decoder.fc_out.bias": (1,)}
    for i in range(NBLK):
        for fc in ("fc_0", "fc_1"):
            keys[f"decoder.blocks.{i}.{fc}.weight"] = (HID, HID, 1)
            keys[f"decoder.blocks.{i}.{fc}.bias"] = (HID,)
    for p in _cbn_prefixes():
        keys[f"{p}.conv_gamma.weight"] = (HID, latent_dim, 1)
        keys[f"{p}.conv_gamma.bias"] = (HID,)
        keys[f"{p}.conv_beta.weight"] = (HID, latent_dim, 1)
        keys[f"{p}.conv_beta.bias"] = (HID,)
        keys[f"{p}.bn.running_mean"] = (HID,)
        keys[f"{p}.bn.running_var"] = (HID,)
        keys[f"{p}.bn.num_batches_tracked"] = ()
    return keys


def pack_decoder(state_dict, latent_dim):
    """Flatten a `ckpt["decoder"]` state_dict into the float32 blob surfd_dec_create() consumes.

    Strict like the reference's load_state_dict(strict=True): missing / unexpected keys raise."""
    exp = expected_keys(latent_dim)
    missing = [k for k in exp if k not in state_dict]
    unexpected = [k for k in state_dict if k not in exp]
    if missing or unexpected:
        raise RuntimeError(f"Error(s) in loading state_dict for CbnDecoder: missing {missing}, unexpected {unexpected}")
    for k, shp in exp.items():
        if tuple(state_dict[k].shape) != shp:
            raise RuntimeError(f"size mismatch for {k}: {tuple(state_dict[k].shape)} vs {shp}")
    f = lambda k: state_dict[k].detach().to(torch.float32).cpu()
    parts = []
    wp = torch.zeros(HID, ENC)
    wp[:, :63] = f("decoder.fc_p.weight")[:, :, 0]
    parts += [wp.reshape(-1), f("decoder.fc_p.bias")]
    for i in range(NBLK):
        for fc in ("fc_0", "fc_1"):
            parts += [f(f"decoder.blocks.{i}.{fc}.weight")[:, :, 0].reshape(-1), f(f"decoder.blocks.{i}.{fc}.bias")]
    bout = torch.zeros(4)
    bout[0] = f("decoder.fc_out.bias")[0]
    parts += [f("decoder.fc_out.weight").reshape(-1), bout]
    for p in _cbn_prefixes():
        parts += [f(f"{p}.conv_gamma.weight")[:, :, 0].reshape(-1), f(f"{p}.conv_gamma.bias"),
                  f(f"{p}.conv_beta.weight")[:, :, 0].reshape(-1), f(f"{p}.conv_beta.bias"),
                  f(f"{p}.bn.running_mean"), f(f"{p}.bn.running_var")]
    blob = torch.cat([p.reshape(-1) for p in parts]).contiguous()
    return blob


class UdfDecoder:
    """Device-resident decoder; one instance per GPU process.  `set_latent` plays the role of closing
    `udf_func` over one shape's latent code."""

    def __init__(self, state_dict, latent_dim, device="cuda", max_chunk_points=0, packed=None):
        lib = _lib.load()
        self.lib = lib
        self.latent_dim = int(latent_dim)
        self.device = torch.device(device)
        blob = packed if packed is not None else pack_decoder(state_dict, latent_dim)
        assert blob.numel() == lib.surfd_dec_packed_floats(self.latent_dim)
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            on_dev = 1 if blob.is_cuda else 0
            _lib.check(lib.surfd_dec_create(_lib.ptr(blob), blob.numel(), self.latent_dim, on_dev, int(max_chunk_points),
                                            ctypes.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self.lib.surfd_dec_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_precision(self, mode):
        _lib.check(self.lib.surfd_dec_set_precision(self._h, int(mode)))

    def set_chain(self, on):
        """TF32 mode: True (default) = the ten 512x512 layers of a pass in ONE launch (tc_chain_kernel: every CTA walks its own
        128-row panels through all layers -- no launch gaps, no grid barrier); False = one launch per layer.  Bit-identical results."""
        _lib.check(self.lib.surfd_dec_set_chain(self._h, 1 if on else 0))

    def set_sm_budget(self, n_sms):
        """CTAs of the persistent tensor-core GEMM (0 = all SMs); see surfd_dec_set_sm_budget"""
        _lib.check(self.lib.surfd_dec_set_sm_budget(self._h, int(n_sms)))

    @property
    def num_sms(self):
        return int(self.lib.surfd_dec_num_sms(self._h))

    def set_latent(self, lat):
        lat = lat.detach().reshape(-1).to(self.device, torch.float32).contiguous()
        assert lat.numel() == self.latent_dim
        self._lat = lat
        _lib.check(self.lib.surfd_dec_set_latent(self._h, _lib.ptr(lat), _lib.stream_ptr()))

    def query(self, pts, want_grad=False):
        """udf [M] (and -normalize(d udf/dx) [M,3]) at points [M,3] (meshudf.py:209-251 semantics)."""
        pts = pts.detach().to(self.device, torch.float32).contiguous()
        M = pts.shape[0]
        udf = torch.empty(M, device=self.device, dtype=torch.float32)
        grad = torch.empty(M, 3, device=self.device, dtype=torch.float32) if want_grad else None
        _lib.check(self.lib.surfd_udf_query(self._h, _lib.ptr(pts), M, _lib.ptr(udf), _lib.ptr(grad), _lib.stream_ptr()))
        return (udf, grad) if want_grad else udf

    def logits(self, pts):
        """CbnDecoder.forward(encode(pts), lat) -> logits [M] (before the sigmoid of udf_func)"""
        pts = pts.detach().to(self.device, torch.float32).contiguous()
        M = pts.shape[0]
        out = torch.empty(M, device=self.device, dtype=torch.float32)
        _lib.check(self.lib.surfd_dec_logits(self._h, _lib.ptr(pts), M, _lib.ptr(out), _lib.stream_ptr()))
        return out

    def lattice(self, N, use_fast_grid_filler=True, max_dist=0.1, grads=True):
        """(udf [N,N,N], grads [N,N,N,3], counts) -- GridFiller.fill_grid or get_udf_and_grads.  grads=False: udf only
        (utils.GridFiller.fill_grid, utils/utils.py:252-339), the second item is None."""
        udf = torch.empty(N, N, N, device=self.device, dtype=torch.float32)
        grad = torch.empty(N, N, N, 3, device=self.device, dtype=torch.float32) if grads else None
        counts = (ctypes.c_int64 * 2)()
        _lib.check(self.lib.surfd_udf_lattice(self._h, int(N), 1 if use_fast_grid_filler else 0, float(max_dist),
                                              _lib.ptr(udf), _lib.ptr(grad) if grads else None, counts, _lib.stream_ptr()))
        return udf, grad, (int(counts[0]), int(counts[1]))

    def time_layer(self, iters=20, points=None):
        """(ms per launch, points per launch) of the dominant kernel, CUDA-event timed on the current stream"""
        M = int(self.lib.surfd_dec_chunk_points(self._h)) if points is None else int(points)
        ms = ctypes.c_float()
        _lib.check(self.lib.surfd_dec_time_layer(self._h, M, int(iters), ctypes.byref(ms), _lib.stream_ptr()))
        return float(ms.value), M

    def profile(self, on):
        """on=True: start bracketing every layer GEMM of the real chain with CUDA events; on=False: stop and return
        (launches, point rows, summed ms)"""
        if on:
            _lib.check(self.lib.surfd_dec_profile(self._h, 1, None, None, None))
            return None
        n, pts, ms = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_double()
        _lib.check(self.lib.surfd_dec_profile(self._h, 0, ctypes.byref(n), ctypes.byref(pts), ctypes.byref(ms)))
        return int(n.value), int(pts.value), float(ms.value)

    def debug_layer(self, A, blk=0, mode=0):
        """one hidden layer over A [M,512] with the FFMA (mode 0) or tcgen05 (mode 1) kernel -- test hook"""
        A = A.to(self.device, torch.float32).contiguous()
        out = torch.empty_like(A)
        _lib.check(self.lib.surfd_dec_debug_layer(self._h, _lib.ptr(A), A.shape[0], int(blk), int(mode), _lib.ptr(out), _lib.stream_ptr()))
        return out

    def face_filter(self, verts64, faces, N):
        """keep mask [F] (uint8) of meshudf.py:356-379."""
        verts64 = verts64.to(self.device, torch.float64).contiguous()
        faces = faces.to(self.device, torch.int32).contiguous()
        F = faces.shape[0]
        keep = torch.empty(F, device=self.device, dtype=torch.uint8)
        _lib.check(self.lib.surfd_face_filter(self._h, _lib.ptr(verts64), int(verts64.shape[0]), _lib.ptr(faces), F, int(N), _lib.ptr(keep),
                                              _lib.stream_ptr()))
        return keep
