#!/usr/bin/env python
"""Target for an ncu launch list of the denoiser: 3 model evaluations at B=8, L=32 (the 2nd/3rd are warm)."""
import os, sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch
sys.path.insert(0, ".")
from surfd_b200 import synth, unet as U
L, B = 32, int(sys.argv[1]) if len(sys.argv) > 1 else 8
net = U.UNetSampler(synth.synth_mdm(L), L, max_batch=B)
g = torch.Generator().manual_seed(0)
x = torch.randn(B, 1, L, generator=g); t = torch.full((B,), 500)
for _ in range(3):
    net.forward(x, t)
torch.cuda.synchronize()
print("ok")
