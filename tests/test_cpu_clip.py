"""Conditioning encoders (surfd_b200/clip_encoder.py) against golden vectors written by the reference's own
vendored CLIP classes, tokenizer and image helpers on a seeded ViT-B/32-shaped checkpoint (tests/golden/make_golden_clip.py).
torch on CPU -- the same code runs on the device in the CLI (tests/test_gpu_zzz_frontends.py)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from surfd_b200 import synth
from surfd_b200.clip_encoder import ClipEncoder, Tokenizer, find_vocab, image_condition, mask2bbox, sketch_condition

PROMPTS = ["a chair", "A round table.  With four LEGS &amp; a glass top!", "an armchair with a very high back, wooden legs and a red cushion " * 6,
           "it's the sofa's 3 seats", "a lamp"]


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "clip_vitb32.npz"))


@pytest.fixture(scope="module")
def enc():
    return ClipEncoder(synth.synth_clip(77), device="cpu")


def _photo():
    H, W = 180, 240
    yy, xx = np.mgrid[0:H, 0:W]
    photo = np.stack([(xx * 255 // (W - 1)), (yy * 255 // (H - 1)), ((xx + 2 * yy) % 256)], -1).astype(np.uint8)
    mask = (((xx - 170) / 60.0) ** 2 + ((yy - 60) / 50.0) ** 2 <= 1.0)
    return photo, mask


def test_configuration_is_read_off_the_checkpoint(enc):
    assert (enc.vision_width, enc.vision_layers, enc.patch, enc.grid, enc.image_resolution, enc.vision_heads) == (768, 12, 32, 7, 224, 12)
    assert (enc.text_width, enc.text_layers, enc.text_heads, enc.context_length, enc.vocab_size, enc.embed_dim) == (512, 12, 8, 77, 49408, 512)


def test_image_and_text_embeddings_match_the_reference_classes(enc, gold):
    images = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(2024))
    img = enc.encode_image(images).numpy()
    txt = enc.encode_text(torch.from_numpy(gold["tokens"])).numpy()
    assert img.shape == (2, 512) and txt.shape == (5, 512)
    # fp32 on both sides; the embeddings have |x| ~ 0.8 (measured here: 4e-6 image, 1e-6 text -- summation order only)
    assert np.abs(img - gold["img_emb"]).max() < 5e-5, np.abs(img - gold["img_emb"]).max()
    assert np.abs(txt - gold["txt_emb"]).max() < 5e-5, np.abs(txt - gold["txt_emb"]).max()
    # the text feature is taken at the end-of-text position: padding after it must not matter, tokens before it must
    t = torch.from_numpy(gold["tokens"]).clone()
    assert np.abs(enc.encode_text(t[:1]).numpy() - txt[:1]).max() < 1e-5                     # batch independence
    t2 = t[:1].clone(); t2[0, 1] = 1000
    assert np.abs(enc.encode_text(t2).numpy() - txt[:1]).max() > 1e-2


def test_errors_mirror_the_reference(enc):
    with pytest.raises(ValueError):
        enc.encode_image(torch.zeros(1, 3, 128, 128))
    with pytest.raises(ValueError):
        enc.encode_text(torch.zeros(1, 76, dtype=torch.int64))
    with pytest.raises(IndexError):
        enc.encode_text(torch.full((1, 77), 49408, dtype=torch.int64))
    with pytest.raises(NotImplementedError):
        ClipEncoder({"visual.layer1.0.conv1.weight": torch.zeros(1)}, device="cpu")
    sd = synth.synth_clip(1, vision=(128, 2, 32, 2), text=(64, 2), embed_dim=32)
    del sd["transformer.resblocks.1.mlp.c_proj.bias"]
    with pytest.raises(KeyError):
        ClipEncoder(sd, device="cpu")


def test_small_configuration_runs(enc):
    sd = synth.synth_clip(5, vision=(128, 2, 32, 2), text=(64, 2), embed_dim=32)
    e = ClipEncoder(sd, device="cpu")
    assert (e.image_resolution, e.vision_heads, e.text_heads) == (64, 2, 1)
    assert e.encode_image(torch.randn(3, 3, 64, 64)).shape == (3, 32)
    tok = torch.zeros(2, 77, dtype=torch.int64); tok[:, 0] = 49406; tok[:, 1] = 320; tok[:, 2] = 49407
    out = e.encode_text(tok)
    assert out.shape == (2, 32) and torch.equal(out[0], out[1])


def _vocab():
    for cand in (os.environ.get("SURFD_CLIP_VOCAB"), "/root/reference/CLIP/clip/bpe_simple_vocab_16e6.txt.gz"):
        if cand and os.path.exists(cand):
            return cand
    try:
        return find_vocab()
    except FileNotFoundError:
        return None


@pytest.mark.skipif(_vocab() is None, reason="the BPE merges file of a CLIP install is not on this machine (it is data of the "
                                             "upstream package, not shipped here)")
def test_tokenizer_matches_the_reference_tokenizer(gold):
    tok = Tokenizer(_vocab())
    assert (tok.sot, tok.eot, len(tok.encoder)) == (49406, 49407, 49408)
    got = tok.tokenize(PROMPTS, truncate=True)
    assert got.dtype == torch.int32 and tuple(got.shape) == (5, 77)
    assert np.array_equal(got.numpy(), gold["tokens"])
    assert int(got[2, 76]) == 49407 and int((got[2] == 0).sum()) == 0                         # truncated: last token forced to EOT
    with pytest.raises(RuntimeError, match="too long"):
        tok.tokenize(PROMPTS[2])
    assert tok.tokenize("a chair").shape == (1, 77)


def test_image_preparation_matches_the_reference_helpers(enc, gold, tmp_path):
    from PIL import Image
    photo, mask = _photo()
    assert tuple(int(v) for v in mask2bbox(mask)) == tuple(int(v) for v in gold["bbox"])
    ip, mp = str(tmp_path / "photo.png"), str(tmp_path / "mask.png")
    Image.fromarray(photo).save(ip)
    Image.fromarray((mask * 255).astype(np.uint8)).save(mp)
    x = image_condition(ip, mp)
    assert tuple(x.shape) == (1, 3, 224, 224)
    assert np.abs(x[0, :, ::8, ::8].numpy() - gold["prepared_sub"]).max() < 1e-5
    assert np.abs(x[0].mean((1, 2)).numpy() - gold["prepared_mean"]).max() < 1e-5
    assert np.abs(enc.encode_image(x).numpy() - gold["prep_emb"]).max() < 5e-5
    # the sketch script's transform: aspect-preserving bicubic resize + centre crop, 3 channels from a grey image
    sp = str(tmp_path / "sketch.png")
    Image.fromarray(photo[:, :, 0]).save(sp)
    s = sketch_condition(sp)
    assert tuple(s.shape) == (1, 3, 224, 224)
    un = s[0] * torch.tensor([0.26862954, 0.26130258, 0.27577711]).view(3, 1, 1) + torch.tensor([0.48145466, 0.4578275, 0.40821073]).view(3, 1, 1)
    assert float((un[0] - un[1]).abs().max()) < 1e-6 and 0.0 <= float(un.min()) and float(un.max()) <= 1.0 + 1e-6


def test_cli_conditioning_step(tmp_path, gold):
    """surfd_b200.cli._context: checkpoint file (plain state dict, as torch.save writes it) -> tokenizer / image preparation ->
    encoder, the [B,512] tensor the sampler takes; --context_path bypass; missing checkpoint fails loudly"""
    from PIL import Image
    from surfd_b200 import cli
    ck = str(tmp_path / "ViT-B-32.pt")
    torch.save(synth.synth_clip(9, vision=(128, 2, 32, 7), text=(64, 2)), ck)
    base = ["--model_path", "m.pt", "--ae_dir", "ae.pt", "--output_dir", str(tmp_path), "--clip_path", ck]
    photo, mask = _photo()
    ip, mp = str(tmp_path / "photo.png"), str(tmp_path / "mask.png")
    Image.fromarray(photo).save(ip)
    Image.fromarray((mask * 255).astype(np.uint8)).save(mp)
    a = cli.generate_args(base + ["--cond_mode", "img", "--image_path", ip, "--mask_path", mp])
    c = cli._context(a, "image", 3, "cpu")
    assert tuple(c.shape) == (3, 512) and c.dtype == torch.float32 and torch.equal(c[0], c[2]) and float(c.abs().mean()) > 0.1
    a = cli.generate_args(base + ["--cond_mode", "sketch", "--sketch_path", ip])
    s = cli._context(a, "sketch", 1, "cpu")
    assert tuple(s.shape) == (1, 512) and not torch.allclose(s[0], c[0])
    if _vocab() is not None:
        a = cli.generate_args(base + ["--cond_mode", "text", "--prompt", "a chair", "--clip_vocab", _vocab()])
        t = cli._context(a, "text", 2, "cpu")
        from surfd_b200.clip_encoder import ClipEncoder as CE
        want = CE.from_file(ck, "cpu").encode_text(torch.from_numpy(gold["tokens"][:1]))
        assert tuple(t.shape) == (2, 512) and torch.allclose(t[0], want[0], atol=1e-6) and torch.equal(t[0], t[1])
    emb = str(tmp_path / "ctx.pt")
    torch.save(torch.arange(512, dtype=torch.float32).reshape(1, 512), emb)
    a = cli.generate_args(["--model_path", "m.pt", "--ae_dir", "ae.pt", "--cond_mode", "text", "--context_path", emb])
    assert tuple(cli._context(a, "text", 4, "cpu").shape) == (4, 512)
    a = cli.generate_args(["--model_path", "m.pt", "--ae_dir", "ae.pt", "--cond_mode", "text", "--prompt", "x", "--clip_path", str(tmp_path / "none.pt")])
    if cli._clip_checkpoint(a) is None:
        with pytest.raises(SystemExit):
            cli._context(a, "text", 1, "cpu")


@pytest.mark.skipif(_vocab() is None, reason="the BPE merges file of a CLIP install is not on this machine")
def test_mdm_encode_text_runs_once_per_prompt_list(tmp_path, monkeypatch):
    """models/mdm.py:86-97: the reference re-encodes the prompt inside every denoiser call; the drop-in MDM encodes a prompt list
    once (through the `clip` drop-in: load + tokenize) and hands the same tensor to every step"""
    from surfd_b200.diffusion import MDM
    import surfd_b200.compat.clip as clip
    ck = str(tmp_path / "ViT-B-32.pt")
    torch.save(synth.synth_clip(9, vision=(128, 2, 32, 7), text=(64, 2)), ck)
    monkeypatch.setenv("SURFD_CLIP_PATH", ck)
    monkeypatch.setenv("SURFD_CLIP_VOCAB", _vocab())
    m = MDM(cond_mode="text", clip_version="ViT-B/32")
    m._device = "cpu"                                   # (the encoders are device-agnostic torch code; the sampler is not)
    a = m.encode_text(["a chair", "a lamp"])
    assert tuple(a.shape) == (2, 512) and m.encode_text(["a chair", "a lamp"]) is a
    ctx, lab = m.conditioning({"text": ["a chair", "a lamp"]}, 2)
    assert ctx is a and lab is None
    model, preprocess = clip.load("ViT-B/32", device="cpu", jit=False)
    want = model.encode_text(clip.tokenize(["a chair", "a lamp"], truncate=True)).float()
    assert torch.equal(want, a)
    monkeypatch.delenv("SURFD_CLIP_PATH")
    monkeypatch.setenv("HOME", str(tmp_path))           # no ~/.cache/clip/ViT-B-32.pt either
    m2 = MDM(cond_mode="text")
    m2._device = "cpu"
    with pytest.raises(RuntimeError, match="weights are not shipped"):
        m2.encode_text(["a chair"])


REF = "/root/reference"


def _ref_module(name, rel):
    import importlib.util, sys, types
    sys.modules.setdefault("ftfy", types.SimpleNamespace(fix_text=lambda s: s))      # absent here; identity on the strings below
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "CLIP", "clip")), reason="reference tree not present (GPU box)")
def test_tokenizer_against_the_live_reference_tokenizer():
    """the reference's own SimpleTokenizer, imported where it lies, on strings that exercise every branch of the pre-tokeniser:
    contractions, digits (one token each), punctuation runs, non-ASCII letters, HTML entities (unescaped twice), whitespace runs,
    the special tokens, upper case"""
    ref = _ref_module("ref_clip_tok_live", "CLIP/clip/simple_tokenizer.py").SimpleTokenizer()
    tok = Tokenizer(os.path.join(REF, "CLIP", "clip", "bpe_simple_vocab_16e6.txt.gz"))
    texts = ["a chair", "A CHAIR!!!", "it's they're we've I'm you'll he'd isn't", "table no. 42 costs 1999.99$", "  many   spaces\t\nand lines ",
             "caf\u00e9 na\u00efve \u00fcber stra\u00dfe", "&amp;lt;b&amp;gt; bold &lt;i&gt;", "<|startoftext|> a lamp <|endoftext|>", "...---???", "x",
             "\u65e5\u672c\u8a9e \u306e \u6905\u5b50", "emoji \U0001F600 chair", "supercalifragilisticexpialidocious antidisestablishmentarianism",
             "a sofa with 3 seats, 2 arm-rests & a foot_stool (dark-grey)"]
    for t in texts:
        assert tok.encode(t) == ref.encode(t), t
    assert tok.encoder == ref.encoder


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "CLIP", "clip")), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("seed", [1, 2])
def test_encoders_against_the_live_reference_model(seed):
    """the reference's own build_model / CLIP on a small seeded configuration (2 + 2 layers), random images and token rows with
    the end-of-text token at different positions"""
    model_py = _ref_module("ref_clip_model_live", "CLIP/clip/model.py")
    sd = synth.synth_clip(seed, vision=(128, 2, 32, 3), text=(128, 2), embed_dim=64)
    ref = model_py.build_model({k: v.clone() for k, v in sd.items()}).float()
    enc = ClipEncoder(sd, device="cpu")
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(3, 3, 96, 96, generator=g)
    tokens = torch.randint(1, 49000, (4, 77), generator=g)
    for i, pos in enumerate((3, 20, 76, 40)):
        tokens[i, pos] = 49407
        tokens[i, pos + 1:] = 0
    with torch.no_grad():
        assert float((ref.encode_image(images) - enc.encode_image(images)).abs().max()) < 2e-5
        assert float((ref.encode_text(tokens) - enc.encode_text(tokens)).abs().max()) < 2e-5
