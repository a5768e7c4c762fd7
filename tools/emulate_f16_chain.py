#!/usr/bin/env python
"""Numerics of the decoder lead in DESIGN.md section 9, emulated on the CPU (torch, no GPU needed): the ten 512x512 layers with
operands rounded the way each candidate kernel would see them, fp32 accumulation, everything else in fp32.

  fp32      : reference arithmetic
  tf32      : activations and weights rounded to TF32 (rna)                         -- the shipped tensor-core mode
  f16       : activations and weights rounded to IEEE half, saturating             -- measured on the GPU in round 2: fast, but wrong
                                                                                      where activations leave the half range
  f16+scale : each activation ROW scaled by an exact power of two so that its largest element sits in [2^14, 2^15), rounded to
              half, the GEMM result multiplied back; each weight matrix scaled the same way (one scale per matrix)

Decoders: the all-random one (noise field, O(1) activations) and the closed-form 'poly' one (the bench's; large activations).
Prints max / mean |udf - udf_fp32| over 20,000 random points (udf range [0, 0.1])."""
import sys
import torch
sys.path.insert(0, ".")
from surfd_b200 import synth


def rna_tf32(x):
    return ((x.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


def to_half(x):
    return x.clamp(-65504.0, 65504.0).half().float()


def row_scale(x):
    m = x.abs().amax(-1, keepdim=True).clamp_min(1e-30)
    return torch.exp2(14.0 - torch.floor(torch.log2(m)))           # largest element -> [2^14, 2^15)


def matmul(a, W, mode):
    if mode == "fp32":
        return a @ W.t()
    if mode == "tf32":
        return rna_tf32(a.contiguous()) @ rna_tf32(W.contiguous()).t()
    if mode == "f16":
        return to_half(a) @ to_half(W).t()
    sa = row_scale(a)
    sw = row_scale(W.reshape(1, -1)).reshape(())
    return (to_half(a * sa) @ to_half(W * sw).t()) / (sa * sw)


def encode(p):
    outs = [p]
    for j in range(10):
        outs += [torch.sin(p * 2.0 ** j), torch.cos(p * 2.0 ** j)]
    return torch.cat(outs, -1)


def forward(sd, lat, pts, mode):
    w = lambda k: sd[k].float()

    def cbn(prefix, x):
        gamma = w(prefix + ".conv_gamma.weight")[:, :, 0] @ lat + w(prefix + ".conv_gamma.bias")
        beta = w(prefix + ".conv_beta.weight")[:, :, 0] @ lat + w(prefix + ".conv_beta.bias")
        inv = 1.0 / torch.sqrt(w(prefix + ".bn.running_var") + 1e-5)
        return gamma * (x - w(prefix + ".bn.running_mean")) * inv + beta

    net = encode(pts) @ w("decoder.fc_p.weight")[:, :, 0].t() + w("decoder.fc_p.bias")      # K = 63: fp32 FFMA in every mode
    amax = 0.0
    for i in range(5):
        pre = f"decoder.blocks.{i}"
        a0 = torch.relu(cbn(pre + ".bn_0", net)); amax = max(amax, float(a0.max()))
        h = matmul(a0, w(pre + ".fc_0.weight")[:, :, 0], mode) + w(pre + ".fc_0.bias")
        a1 = torch.relu(cbn(pre + ".bn_1", h)); amax = max(amax, float(a1.max()))
        net = net + matmul(a1, w(pre + ".fc_1.weight")[:, :, 0], mode) + w(pre + ".fc_1.bias")
    af = torch.relu(cbn("decoder.bn", net))
    logit = af @ w("decoder.fc_out.weight")[0, :, 0] + w("decoder.fc_out.bias")[0]
    return (1 - torch.sigmoid(logit)) * 0.1, amax


def main():
    torch.manual_seed(0)
    L = 32
    pts = torch.rand(20000, 3) * 2 - 1
    for name, ck in (("random", synth.synth_ae_rand(L, 4321)), ("poly", synth.synth_ae_poly(L))):
        sd = ck["decoder"]
        lat = 0.7 * torch.randn(L, generator=torch.Generator().manual_seed(1))
        with torch.no_grad():
            ref, amax = forward(sd, lat, pts, "fp32")
            print(f"{name:7s} decoder: largest activation entering a 512x512 layer {amax:.3g}; largest |weight| "
                  f"{max(float(v.abs().max()) for k, v in sd.items() if 'fc_0.weight' in k or 'fc_1.weight' in k):.3g}")
            for mode in ("tf32", "f16", "f16+scale"):
                u, _ = forward(sd, lat, pts, mode)
                d = (u - ref).abs()
                print(f"    {mode:10s} max |d udf| {float(d.max()):.2e}   mean {float(d.mean()):.2e}   99.9th pct {float(d.quantile(0.999)):.2e}")


if __name__ == "__main__":
    main()
