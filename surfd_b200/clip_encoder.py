"""The conditioning encoders in front of the path: CLIP ViT-B/32 image / text encoders, the BPE tokenizer
and the scripts' image preparation, producing the [B, 512] `context` the sampler takes.

Reference (vendored openai/CLIP + the scripts' glue):
  CLIP/clip/model.py:157-240    LayerNorm (fp32), QuickGELU, ResidualAttentionBlock (nn.MultiheadAttention), Transformer,
                                VisionTransformer.forward
  CLIP/clip/model.py:343-358    CLIP.encode_image / encode_text (EOT feature = position of the largest token id)
  CLIP/clip/model.py:399-436    build_model: the configuration is read off the state dict's shapes
  CLIP/clip/simple_tokenizer.py byte-level BPE (49,152 merges file, <|startoftext|> / <|endoftext|>)
  CLIP/clip/clip.py:195-237     tokenize(texts, context_length=77, truncate)
  models/mdm.py:86-97           encode_text(raw_text) = clip_model.encode_text(clip.tokenize(raw_text, truncate=True)).float(),
                                re-run inside EVERY denoiser call (1000 x per generation); here it runs once
  sample/generate_image.py:92-115, data_loaders/dataset.py:19-94   mask2bbox, crop_square, _transform_rgb(224)
  sample/generate_sketch.py:30-37,74-82                           _transform(224) on the sketch image

The scripts load CLIP with `clip.load('ViT-B/32', device='cpu', jit=False)`, i.e. in fp32, and run the image encoder on the
host; here the weights live on the device in fp32 and both encoders run there (one pass per generation: stream-ordered torch
GEMM / attention calls, not a hot path -- 4.4 GFLOP per image, 2.9 per prompt).  Weights are not shipped (no network):
`ClipEncoder.from_file` reads an OpenAI CLIP checkpoint (TorchScript archive or plain state dict), the tokenizer needs the
merges file of a CLIP install (`bpe_simple_vocab_16e6.txt.gz`).  Only the ViT variants are built (the scripts use ViT-B/32).
Parity: golden vectors written by the reference's own classes on seeded synthetic weights (tests/golden/make_golden_clip.py).
"""
import gzip
import html
import os
from functools import lru_cache

import torch
import torch.nn.functional as F

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def _ln(x, w, b):
    return F.layer_norm(x.float(), (x.shape[-1],), w, b, 1e-5)


class ClipEncoder:
    """encode_image([B,3,R,R]) / encode_text(int tokens [B,77]) -> [B, embed_dim] fp32 on `device`."""

    def __init__(self, state_dict, device="cuda"):
        sd = state_dict
        if "visual.proj" not in sd:
            raise NotImplementedError("only the ViT image towers are built (the scripts use ViT-B/32); this checkpoint is a ResNet CLIP")
        self.device = torch.device(device)
        # configuration from the shapes, like build_model (model.py:399-420)
        self.vision_width = sd["visual.conv1.weight"].shape[0]
        self.vision_layers = len([k for k in sd if k.startswith("visual.") and k.endswith(".attn.in_proj_weight")])
        self.patch = sd["visual.conv1.weight"].shape[-1]
        self.grid = round((sd["visual.positional_embedding"].shape[0] - 1) ** 0.5)
        self.image_resolution = self.patch * self.grid
        self.embed_dim = sd["text_projection"].shape[1]
        self.context_length = sd["positional_embedding"].shape[0]
        self.vocab_size = sd["token_embedding.weight"].shape[0]
        self.text_width = sd["ln_final.weight"].shape[0]
        self.text_layers = len(set(k.split(".")[2] for k in sd if k.startswith("transformer.resblocks")))
        self.vision_heads = self.vision_width // 64
        self.text_heads = self.text_width // 64
        skip = ("input_resolution", "context_length", "vocab_size", "logit_scale")
        self.w = {k: v.detach().to(self.device, torch.float32).contiguous() for k, v in sd.items() if k not in skip}
        for tower, n in (("visual.transformer", self.vision_layers), ("transformer", self.text_layers)):
            for i in range(n):
                for leaf in ("ln_1.weight", "ln_1.bias", "attn.in_proj_weight", "attn.in_proj_bias", "attn.out_proj.weight",
                             "attn.out_proj.bias", "ln_2.weight", "ln_2.bias", "mlp.c_fc.weight", "mlp.c_fc.bias",
                             "mlp.c_proj.weight", "mlp.c_proj.bias"):
                    if f"{tower}.resblocks.{i}.{leaf}" not in self.w:
                        raise KeyError(f"CLIP checkpoint lacks {tower}.resblocks.{i}.{leaf}")
        # the patch convolution as one GEMM: [width, 3 * patch * patch]
        self.w_patch = self.w["visual.conv1.weight"].reshape(self.vision_width, -1)
        self.causal = torch.full((self.context_length, self.context_length), float("-inf"), device=self.device).triu_(1)

    @classmethod
    def from_file(cls, path, device="cuda"):
        """an OpenAI checkpoint as `clip.load` reads it (clip.py:128-140): the published files are TorchScript archives, a
        plain state dict (torch.save) is accepted as well"""
        try:
            sd = torch.load(path, map_location="cpu", weights_only=True)
            if isinstance(sd, dict) and "state_dict" in sd:
                sd = sd["state_dict"]
        except Exception:
            sd = torch.jit.load(path, map_location="cpu").state_dict()
        return cls(sd, device)

    # ---- transformer (model.py:171-203) ----
    def _blocks(self, x, tower, layers, heads, mask):
        """x [B, T, W] -> [B, T, W]; pre-LN residual blocks, QuickGELU MLP"""
        B, T, W = x.shape
        hd = W // heads
        for i in range(layers):
            p = f"{tower}.resblocks.{i}."
            h = _ln(x, self.w[p + "ln_1.weight"], self.w[p + "ln_1.bias"])
            qkv = F.linear(h, self.w[p + "attn.in_proj_weight"], self.w[p + "attn.in_proj_bias"])
            q, k, v = (t.reshape(B, T, heads, hd).transpose(1, 2) for t in qkv.chunk(3, dim=-1))
            att = (q * hd ** -0.5) @ k.transpose(-1, -2)                                   # [B, heads, T, T]; T <= 77
            if mask is not None:
                att = att + mask
            a = (torch.softmax(att, dim=-1) @ v).transpose(1, 2).reshape(B, T, W)
            x = x + F.linear(a, self.w[p + "attn.out_proj.weight"], self.w[p + "attn.out_proj.bias"])
            h = _ln(x, self.w[p + "ln_2.weight"], self.w[p + "ln_2.bias"])
            h = F.linear(h, self.w[p + "mlp.c_fc.weight"], self.w[p + "mlp.c_fc.bias"])
            h = h * torch.sigmoid(1.702 * h)
            x = x + F.linear(h, self.w[p + "mlp.c_proj.weight"], self.w[p + "mlp.c_proj.bias"])
        return x

    @torch.no_grad()
    def encode_image(self, image):
        """VisionTransformer.forward (model.py:223-240)"""
        x = image.detach().to(self.device, torch.float32)
        R, g, p = self.image_resolution, self.grid, self.patch
        if x.dim() != 4 or tuple(x.shape[1:]) != (3, R, R):
            raise ValueError(f"image batch must have shape (B, 3, {R}, {R}), got {tuple(x.shape)}")
        B = x.shape[0]
        patches = x.reshape(B, 3, g, p, g, p).permute(0, 2, 4, 1, 3, 5).reshape(B, g * g, 3 * p * p)
        x = patches @ self.w_patch.t()                                                     # conv1, stride = kernel, no bias
        x = torch.cat([self.w["visual.class_embedding"].expand(B, 1, -1), x], 1) + self.w["visual.positional_embedding"]
        x = _ln(x, self.w["visual.ln_pre.weight"], self.w["visual.ln_pre.bias"])
        x = self._blocks(x, "visual.transformer", self.vision_layers, self.vision_heads, None)
        x = _ln(x[:, 0], self.w["visual.ln_post.weight"], self.w["visual.ln_post.bias"])
        return x @ self.w["visual.proj"]

    @torch.no_grad()
    def encode_text(self, tokens):
        """CLIP.encode_text (model.py:346-358)"""
        t = tokens.detach().to(self.device, torch.int64)
        if t.dim() != 2 or t.shape[1] != self.context_length:
            raise ValueError(f"tokens must have shape (B, {self.context_length}), got {tuple(t.shape)}")
        if int(t.min()) < 0 or int(t.max()) >= self.vocab_size:
            raise IndexError("index out of range in self: token id outside the vocabulary")
        x = self.w["token_embedding.weight"][t] + self.w["positional_embedding"]
        x = self._blocks(x, "transformer", self.text_layers, self.text_heads, self.causal)
        x = _ln(x, self.w["ln_final.weight"], self.w["ln_final.bias"])
        return x[torch.arange(x.shape[0], device=self.device), t.argmax(dim=-1)] @ self.w["text_projection"]


# ---- tokenizer (simple_tokenizer.py) --------------------------------------------------------------------------------------

@lru_cache()
def _bytes_to_unicode():
    """the reversible byte <-> printable-character table of the BPE (simple_tokenizer.py:15-36)"""
    keep = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAC + 1)) + list(range(0xAE, 0xFF + 1))
    table, extra = {}, 0
    for b in keep:
        table[b] = chr(b)
    for b in range(256):
        if b not in table:
            table[b] = chr(256 + extra)
            extra += 1
    # vocabulary order is: the kept bytes in `keep` order, then the remapped ones in byte order
    order = keep + [b for b in range(256) if b not in keep]
    return table, [table[b] for b in order]


def find_vocab(hint=None):
    """merges file of a CLIP install: explicit path, $SURFD_CLIP_VOCAB, next to `hint`, or an importable `clip` package"""
    cands = [hint if hint and hint.endswith(".gz") else None, os.environ.get("SURFD_CLIP_VOCAB")]
    if hint and not hint.endswith(".gz"):
        cands.append(os.path.join(os.path.dirname(os.path.abspath(hint)), "bpe_simple_vocab_16e6.txt.gz"))
    try:
        import clip as _clip
        cands.append(os.path.join(os.path.dirname(os.path.abspath(_clip.__file__)), "bpe_simple_vocab_16e6.txt.gz"))
    except Exception:
        pass
    for c in cands:
        if c and os.path.exists(c):
            return c
    raise FileNotFoundError("CLIP merges file bpe_simple_vocab_16e6.txt.gz not found: pass --clip_vocab or set SURFD_CLIP_VOCAB")


class Tokenizer:
    def __init__(self, bpe_path):
        import regex
        self.byte_table, base = _bytes_to_unicode()
        lines = gzip.open(bpe_path).read().decode("utf-8").split("\n")
        merges = [tuple(m.split()) for m in lines[1:49152 - 256 - 2 + 1]]
        vocab = base + [c + "</w>" for c in base] + ["".join(m) for m in merges] + ["<|startoftext|>", "<|endoftext|>"]
        self.encoder = {tok: i for i, tok in enumerate(vocab)}
        self.ranks = {m: i for i, m in enumerate(merges)}
        self.cache = {"<|startoftext|>": ("<|startoftext|>",), "<|endoftext|>": ("<|endoftext|>",)}
        self.pat = regex.compile(r"""<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+""",
                                 regex.IGNORECASE)
        self._ws = regex.compile(r"\s+")
        self.sot, self.eot = self.encoder["<|startoftext|>"], self.encoder["<|endoftext|>"]

    def _bpe(self, token):
        """greedy lowest-rank pair merging of one pre-token; returns the tuple of sub-word symbols"""
        if token in self.cache:
            return self.cache[token]
        word = list(token[:-1]) + [token[-1] + "</w>"]
        while len(word) > 1:
            best, rank = None, None
            for pair in zip(word[:-1], word[1:]):
                r = self.ranks.get(pair)
                if r is not None and (rank is None or r < rank):
                    best, rank = pair, r
            if best is None:
                break
            merged, i = [], 0
            while i < len(word):
                if i + 1 < len(word) and word[i] == best[0] and word[i + 1] == best[1]:
                    merged.append(best[0] + best[1]); i += 2
                else:
                    merged.append(word[i]); i += 1
            word = merged
        out = tuple(word)
        self.cache[token] = out
        return out

    def encode(self, text):
        try:                                      # basic_clean: ftfy.fix_text when ftfy exists (identity on clean ASCII text)
            import ftfy
            text = ftfy.fix_text(text)
        except ImportError:
            pass
        text = html.unescape(html.unescape(text)).strip()
        text = self._ws.sub(" ", text).strip().lower()
        ids = []
        for tok in self.pat.findall(text):
            tok = "".join(self.byte_table[b] for b in tok.encode("utf-8"))
            ids.extend(self.encoder[s] for s in self._bpe(tok))
        return ids

    def tokenize(self, texts, context_length=77, truncate=False):
        """clip.tokenize (clip.py:195-237): [sot] + ids + [eot], zero padded; int32 [n, context_length]"""
        if isinstance(texts, str):
            texts = [texts]
        out = torch.zeros(len(texts), context_length, dtype=torch.int32)
        for i, text in enumerate(texts):
            ids = [self.sot] + self.encode(text) + [self.eot]
            if len(ids) > context_length:
                if not truncate:
                    raise RuntimeError(f"Input {text} is too long for context length {context_length}")
                ids = ids[:context_length]
                ids[-1] = self.eot
            out[i, :len(ids)] = torch.tensor(ids, dtype=torch.int32)
        return out


# ---- image preparation of the scripts --------------------------------------------------------------------------------

def mask2bbox(mask):
    """data_loaders/dataset.py:19-26: (x0, y0, x1, y1) of the mask's non-zero extent"""
    import numpy as np
    rows = np.where(np.any(mask, axis=1))[0]
    cols = np.where(np.any(mask, axis=0))[0]
    return cols[0], rows[0], cols[-1], rows[-1]


def crop_square(img, bbox, img_size_h=256, img_size_w=256):
    """data_loaders/dataset.py:29-76 (from Pix2Vox): square crop around the box, edge padding where it leaves the image, PIL
    resize (bicubic, PIL's default) to 256 x 256"""
    import numpy as np
    from PIL import Image
    H, W, _ = img.shape
    x0, y0, x1, y1 = bbox
    side = max(x1 - x0, y1 - y0)
    xm, ym = (x0 + x1) * .5, (y0 + y1) * .5
    xl, xr = int(xm - side * .5), int(xm + side * .5)
    yt, yb = int(ym - side * .5), int(ym + side * .5)
    pl = pr = pt = pb = 0
    if xl < 0:
        pl, xl = -xl, 0
    if xr >= W:
        pr, xr = xr - W + 1, W - 1
    if yt < 0:
        pt, yt = -yt, 0
    if yb >= H:
        pb, yb = yb - H + 1, H - 1
    out = np.pad(img[yt:yb + 1, xl:xr + 1], ((pt, pb), (pl, pr), (0, 0)), mode="edge")
    return Image.fromarray(out).resize((img_size_w, img_size_h))


def image_condition(image_path, mask_path, n_px=224):
    """sample/generate_image.py:92-111: masked image -> square crop -> ToTensor, Normalize, Resize((224, 224)); [1,3,224,224]"""
    import numpy as np
    from PIL import Image
    from torchvision.transforms import Compose, Normalize, Resize, ToTensor
    img_np = np.array(Image.open(image_path).convert("RGB"))
    mask_np = np.array(Image.open(mask_path).convert("1"))
    bbox = list(mask2bbox(mask_np))
    img_clean = (img_np * mask_np[:, :, None]).astype(np.uint8)
    img_clean = crop_square(img_clean, bbox)
    tf = Compose([ToTensor(), Normalize(CLIP_MEAN, CLIP_STD), Resize((n_px, n_px))])      # _transform_rgb (dataset.py:88-93)
    return tf(img_clean).unsqueeze(0)


def sketch_condition(sketch_path, n_px=224):
    """sample/generate_sketch.py:30-37,74-77: Resize(224, bicubic), CenterCrop, RGB, ToTensor, Normalize; [1,3,224,224]"""
    from PIL import Image
    from torchvision.transforms import CenterCrop, Compose, InterpolationMode, Normalize, Resize, ToTensor
    tf = Compose([Resize(n_px, interpolation=InterpolationMode.BICUBIC), CenterCrop(n_px), lambda im: im.convert("RGB"), ToTensor(),
                  Normalize(CLIP_MEAN, CLIP_STD)])
    return tf(Image.open(sketch_path)).unsqueeze(0)
