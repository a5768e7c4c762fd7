#!/usr/bin/env python
"""Target for an `ncu --set full` capture of the persistent sampler kernel: a 4-step reverse process at B=8, L=32."""
import os, sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch
sys.path.insert(0, ".")
from surfd_b200 import synth, unet as U
L, B = 32, int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
net = U.UNetSampler(synth.synth_mdm(L), L, max_batch=B)
S = U.SpacedSchedule(U.cosine_betas(), U.space_timesteps(1000, [n]))
noise = torch.randn(n + 1, B, L, generator=torch.Generator().manual_seed(0)).cuda()
for _ in range(2):
    out = net.sample(S, noise)
torch.cuda.synchronize(); net.status()
print("ok", float(out.abs().max()))
