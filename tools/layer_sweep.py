#!/usr/bin/env python
"""Decoder layer GEMM (tcgen05 TF32 / fp32 FFMA) rate against the points per launch (L2 residency of the activations)."""
import os, sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch
sys.path.insert(0, ".")
from surfd_b200 import synth
from surfd_b200.decoder import UdfDecoder

L = 32
for budget in (148, 140):
    dec = UdfDecoder(synth.synth_ae_poly(L)["decoder"], L, max_chunk_points=148 * 512)
    dec.set_precision(1); dec.set_sm_budget(budget)
    dec.set_latent(torch.randn(L, generator=torch.Generator().manual_seed(0)))
    for tiles_per_cta in (0.5, 1, 2, 3, 4):
        M = int(budget * 128 * tiles_per_cta)
        ms, m = dec.time_layer(30, M)
        print("CTAs %d  points/launch %6d (%.1f MB per activation buffer): %.4f ms  %.1f TFLOP/s" % (budget, m, m * 2048 / 1e6, ms, 2 * m * 512 * 512 / ms / 1e9))
