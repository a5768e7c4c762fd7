"""`import clip` of the reference's scripts (models/mdm.py:3, sample/generate_*.py), backed by surfd_b200.clip_encoder:
`clip.load('ViT-B/32', device=..., jit=False)` -> (model, preprocess) and `clip.tokenize(texts, truncate=...)`.

The weights are not shipped: `load` takes the checkpoint from $SURFD_CLIP_PATH or ~/.cache/clip/ViT-B-32.pt (where upstream
`clip.load` downloads it, CLIP/clip/clip.py:94-140); `tokenize` needs the merges file ($SURFD_CLIP_VOCAB, next to the
checkpoint, or an installed upstream package)."""
import os

import torch

from ... import clip_encoder as _ce

_tokenizer = None
_last_checkpoint = None


class _Model:
    """the three things the scripts do with the CLIP module: eval(), parameters() (to freeze them), encode_image / encode_text"""

    def __init__(self, enc):
        self.enc = enc

    def eval(self):
        return self

    def float(self):
        return self

    def to(self, device):
        if torch.device(device) != self.enc.device:
            self.enc = _ce.ClipEncoder({k: v for k, v in self.enc.w.items()}, device)
        return self

    def cuda(self):
        return self.to("cuda")

    def parameters(self):
        return iter(())            # frozen by construction (inference-only tensors)

    def encode_image(self, image):
        return self.enc.encode_image(image)

    def encode_text(self, text):
        return self.enc.encode_text(text)


def available_models():
    return ["ViT-B/32"]


def load(name="ViT-B/32", device="cuda" if torch.cuda.is_available() else "cpu", jit=False, download_root=None):
    global _last_checkpoint
    if name not in ("ViT-B/32",) and not os.path.isfile(name):
        raise RuntimeError(f"Model {name} not found; available models = {available_models()}")
    cands = [name if os.path.isfile(name) else None, os.environ.get("SURFD_CLIP_PATH"),
             os.path.join(download_root or os.path.expanduser("~/.cache/clip"), "ViT-B-32.pt")]
    path = next((c for c in cands if c and os.path.exists(c)), None)
    if path is None:
        raise RuntimeError("CLIP ViT-B/32 weights are not shipped and there is no network: set SURFD_CLIP_PATH to ViT-B-32.pt")
    _last_checkpoint = path
    enc = _ce.ClipEncoder.from_file(path, device)
    from torchvision.transforms import CenterCrop, Compose, InterpolationMode, Normalize, Resize, ToTensor
    n_px = enc.image_resolution
    preprocess = Compose([Resize(n_px, interpolation=InterpolationMode.BICUBIC), CenterCrop(n_px), lambda im: im.convert("RGB"), ToTensor(),
                          Normalize(_ce.CLIP_MEAN, _ce.CLIP_STD)])                          # clip.py:79-86
    return _Model(enc), preprocess


def tokenize(texts, context_length=77, truncate=False):
    global _tokenizer
    if _tokenizer is None:
        _tokenizer = _ce.Tokenizer(_ce.find_vocab(_last_checkpoint))
    return _tokenizer.tokenize(texts, context_length, truncate)
