#!/usr/bin/env python
"""Golden vectors for the conditioning encoders (surfd_b200/clip_encoder.py), written by running the REFERENCE's own vendored
CLIP (/root/reference/CLIP/clip/model.py, simple_tokenizer.py) and image helpers (data_loaders/dataset.py) on seeded inputs.
Run in the build container only:  python tests/golden/make_golden_clip.py

The checkpoint is surfd_b200.synth.synth_clip (seeded, full ViT-B/32 shapes; OpenAI weights are not available offline).
`ftfy` is absent here: a one-function stub (fix_text = identity, exact for the ASCII prompts used) lets the reference's
tokenizer import.  Outputs: tests/golden/clip_vitb32.npz (a few KB)."""
import importlib.util, os, sys, types
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True


def load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


PROMPTS = ["a chair", "A round table.  With four LEGS &amp; a glass top!", "an armchair with a very high back, wooden legs and a red cushion " * 6,
           "it's the sofa's 3 seats", "a lamp"]


def inputs():
    g = torch.Generator().manual_seed(2024)
    images = torch.randn(2, 3, 224, 224, generator=g)
    # a synthetic photo + mask for the image-preparation helpers: smooth colour gradients, an off-centre elliptical mask
    H, W = 180, 240
    yy, xx = np.mgrid[0:H, 0:W]
    photo = np.stack([(xx * 255 // (W - 1)), (yy * 255 // (H - 1)), ((xx + 2 * yy) % 256)], -1).astype(np.uint8)
    mask = (((xx - 170) / 60.0) ** 2 + ((yy - 60) / 50.0) ** 2 <= 1.0)
    return images, photo, mask


def main():
    from surfd_b200.synth import synth_clip
    sys.modules["ftfy"] = types.SimpleNamespace(fix_text=lambda s: s)
    model_py = load("ref_clip_model", os.path.join(REF, "CLIP", "clip", "model.py"))
    tok_py = load("ref_clip_tok", os.path.join(REF, "CLIP", "clip", "simple_tokenizer.py"))
    sd = synth_clip(77)
    model = model_py.build_model({k: v.clone() for k, v in sd.items()}).float()      # clip.load(device='cpu'): model.float()
    tok = tok_py.SimpleTokenizer()
    sot, eot = tok.encoder["<|startoftext|>"], tok.encoder["<|endoftext|>"]
    tokens = torch.zeros(len(PROMPTS), 77, dtype=torch.int64)
    for i, p in enumerate(PROMPTS):                                                   # clip.tokenize(..., truncate=True)
        ids = [sot] + tok.encode(p) + [eot]
        if len(ids) > 77:
            ids = ids[:77]; ids[-1] = eot
        tokens[i, :len(ids)] = torch.tensor(ids)
    images, photo, mask = inputs()
    with torch.no_grad():
        img_emb = model.encode_image(images).float()
        txt_emb = model.encode_text(tokens).float()
    # image preparation through the reference's helpers (dataset.py imports the training stack; take the three functions by source)
    src = open(os.path.join(REF, "data_loaders", "dataset.py")).read().split("\n")
    ns = {}
    exec("import numpy as np\nfrom PIL import Image\nfrom torchvision.transforms import Compose, Resize, CenterCrop, ToTensor, Normalize\n" +
         "\n".join(src[18:94]), ns)
    x0, y0, x1, y1 = ns["mask2bbox"](mask)
    clean = (photo * mask[:, :, None]).astype(np.uint8)
    prepared = ns["_transform_rgb"](224)(ns["crop_square"](clean, [x0, y0, x1, y1]))
    with torch.no_grad():
        prep_emb = model.encode_image(prepared.unsqueeze(0)).float()
    np.savez_compressed(os.path.join(OUT, "clip_vitb32.npz"), tokens=tokens.numpy().astype(np.int32), img_emb=img_emb.numpy(),
                        txt_emb=txt_emb.numpy(), bbox=np.array([x0, y0, x1, y1]), prepared_sub=prepared[:, ::8, ::8].numpy(),
                        prepared_mean=prepared.mean((1, 2)).numpy(), prep_emb=prep_emb.numpy())
    print("tokens", tokens[:, :8].tolist(), "img", float(img_emb.abs().mean()), "txt", float(txt_emb.abs().mean()), "bbox", (x0, y0, x1, y1))


if __name__ == "__main__":
    main()
