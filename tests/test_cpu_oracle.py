"""CPU suite (-m "not gpu"): the oracles against the committed reference-generated goldens, host logic,
and the C-ABI surface.  Nothing here computes on a GPU."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT, GOLDEN
from oracle import decoder_oracle as O
from oracle import gridfiller_oracle as G
from surfd_b200 import synth
from surfd_b200.decoder import pack_decoder, expected_keys


@pytest.mark.parametrize("L", [32, 64])
def test_decoder_oracle_matches_reference_golden(L):
    g = np.load(os.path.join(GOLDEN, f"decoder_L{L}.npz"))
    sd = synth.synth_ae_rand(L, 4321)["decoder"]
    udf, grads = O.forward(sd, g["lat"][0], g["pts"], want_grad=True)
    assert np.abs(udf - g["udf"]).max() < 1e-6           # fp32 tolerance on udf in [0, 0.1]
    assert np.abs(grads - g["grads"]).max() < 2e-4       # unit vectors; ~1e-3 rad
    assert ((np.abs(grads).sum(-1) == 0) == (np.abs(g["grads"]).sum(-1) == 0)).all()   # identical zero set


def test_poly_checkpoint_encodes_the_polytope_udf():
    L = 32
    sd = synth.synth_ae_poly(L)["decoder"]
    gen = torch.Generator().manual_seed(5)
    lat = torch.randn(L, generator=gen)
    pts = torch.rand(4096, 3, generator=gen) * 2 - 1
    udf = O.forward(sd, lat.numpy(), pts.numpy())
    exact, m = synth.poly_udf(pts, lat)
    assert np.abs(udf - exact.numpy()).max() < 1e-6
    near = np.abs(m.numpy()) < 0.02
    assert np.abs(udf[near] - np.abs(m.numpy()[near])).max() < 2e-4   # PL logit reproduces |m| near the surface


def test_gridfiller_oracle_matches_reference_golden():
    g = np.load(os.path.join(GOLDEN, "gridfiller_poly_N64.npz"))
    sd = synth.synth_ae_poly(32)["decoder"]
    lat = g["lat"][0]

    def chunked(fn, pts, n=32768):
        return np.concatenate([fn(pts[i:i + n]) for i in range(0, len(pts), n)], 0)
    uf = lambda p: chunked(lambda q: O.forward(sd, lat, q), p)
    gf = lambda p: chunked(lambda q: O.forward(sd, lat, q, True)[1], p)
    u, gr, info = G.fill_grid(uf, gf, 64)
    assert info["n_udf"] + info["n_grad"] == int(g["gf_calls"][0])
    assert np.abs(u - g["gf_udf"]).max() < 1e-6
    mask = np.unpackbits(g["gf_gradmask"])[:64 ** 3].astype(bool).reshape(64, 64, 64)
    assert ((np.abs(gr).sum(-1) > 0) == mask).all()
    assert np.abs(gr[mask] - g["gf_grads"].astype(np.float32)).max() < 1e-3   # golden grads stored as fp16


def test_pack_decoder_is_strict_like_load_state_dict():
    sd = synth.synth_ae_rand(32, 1)["decoder"]
    blob = pack_decoder(sd, 32)
    n = 512 * 64 + 512 + 10 * (512 * 512 + 512) + 512 + 4 + 11 * (512 * 32 * 2 + 4 * 512)
    assert blob.numel() == n and blob.dtype == torch.float32
    bad = dict(sd); bad.pop("decoder.fc_out.bias")
    with pytest.raises(RuntimeError):
        pack_decoder(bad, 32)
    bad = dict(sd); bad["decoder.extra"] = torch.zeros(1)
    with pytest.raises(RuntimeError):
        pack_decoder(bad, 32)
    with pytest.raises(RuntimeError):
        pack_decoder(sd, 64)   # wrong latent size -> shape mismatch
    assert len(expected_keys(32)) == 4 + 20 + 11 * 7


def test_library_builds_and_exports_every_declared_symbol():
    from surfd_b200 import build, _lib
    so = build.build()
    assert os.path.exists(so)
    header = open(os.path.join(ROOT, "include", "surfd_b200.h")).read()
    declared = set(re.findall(r"\b(surfd_[a-z0-9_]+)\s*\(", header))
    lib = ctypes.CDLL(so)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/surfd_b200.h but not exported"
    assert declared == set(_lib.PROTOTYPES), "ctypes prototypes and header disagree"
    assert _lib.load().surfd_version() >= 100
    assert _lib.load().surfd_dec_packed_floats(32) == pack_decoder(synth.synth_ae_rand(32, 1)["decoder"], 32).numel()


def test_product_has_no_oracle_import():
    """the product package must never import the oracle (or any CPU fallback)"""
    pkg = os.path.join(ROOT, "surfd_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in re.sub(r'"""[\s\S]*?"""', "", src).replace("# ", ""), os.path.join(dirpath, f)
