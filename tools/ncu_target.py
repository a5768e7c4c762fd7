#!/usr/bin/env python
"""Small target for `ncu --set full` captures: one TF32 decoder query (tc_gemm_kernel), one fp32 query (sgemm_nt_kernel),
one N=256 lattice + marching cubes (classify_kernel, replay_kernel), two sampler steps (conv_gemm_kernel, gn_kernel...)."""
import sys
import torch
sys.path.insert(0, ".")
from surfd_b200 import synth, unet as U
from surfd_b200.decoder import UdfDecoder
from surfd_b200.meshudf import MarchingCubes

L = 32
gen = torch.Generator().manual_seed(0)
lat = torch.randn(L, generator=gen)
pts = (torch.rand(37888 * 2, 3, generator=gen) * 2 - 1).cuda()
sd = synth.synth_ae_poly(L)["decoder"]
tc = UdfDecoder(sd, L); tc.set_precision(1); tc.set_latent(lat)
tc.query(pts)
fp = UdfDecoder(sd, L); fp.set_latent(lat)
fp.query(pts[:37888])
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
u, g, c = tc.lattice(N, True)
mc = MarchingCubes()
mc.classify(u)
net = U.UNetSampler(synth.synth_mdm(L), L, max_batch=8)
net.forward(torch.randn(8, 1, L, generator=gen), torch.full((8,), 500))
torch.cuda.synchronize()
print("ok")
