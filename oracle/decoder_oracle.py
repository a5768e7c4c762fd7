"""ORACLE (test infrastructure; never imported by the product path).

numpy float32 restatement of the reference's decoder query:
  CoordsEncoder.encode                      AutoEncoder/models/coordsenc.py:34-51
  CbnDecoder / DecoderConditionalBatchNorm  AutoEncoder/models/cbndec.py:16-47, 68-82, 99-103, 127-134
  udf_func                                  sample/generate_uncond.py:96-101
  sample_grads                              meshudf/meshudf.py:231-251   (-F.normalize(autograd grad))
Pinned against the reference itself (imported from /root/reference) by tests/golden/make_golden.py; the
committed vectors are tests/golden/decoder_*.npz.
"""
import numpy as np

f32 = np.float32


def encode(p):
    """[M,3] -> [M,63]: cat([x, sin(x*1), cos(x*1), sin(x*2), ... cos(x*512)], -1)"""
    p = np.asarray(p, f32)
    outs = [p]
    for j in range(10):
        f = f32(2.0 ** j)
        outs.append(np.sin(p * f, dtype=f32))
        outs.append(np.cos(p * f, dtype=f32))
    return np.concatenate(outs, -1).astype(f32)


def _w(sd, k):
    a = sd[k]
    a = a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)
    return a.astype(f32)


def _cbn(sd, prefix, x, lat):
    """gamma(z) * BN_eval(x) + beta(z);  BN_eval(x) = (x - mean) / sqrt(var + 1e-5)"""
    gamma = _w(sd, prefix + ".conv_gamma.weight")[:, :, 0] @ lat + _w(sd, prefix + ".conv_gamma.bias")
    beta = _w(sd, prefix + ".conv_beta.weight")[:, :, 0] @ lat + _w(sd, prefix + ".conv_beta.bias")
    mean = _w(sd, prefix + ".bn.running_mean")
    var = _w(sd, prefix + ".bn.running_var")
    inv = (f32(1.0) / np.sqrt(var + f32(1e-5))).astype(f32)
    xn = (x - mean) * inv
    return (gamma * xn + beta).astype(f32), (gamma * inv).astype(f32)


def forward(sd, lat, pts, want_grad=False):
    """udf [M] (and -normalize(d udf / d pts) [M,3]) for one latent [L]."""
    lat = np.asarray(lat, f32).reshape(-1)
    pts = np.asarray(pts, f32)
    e = encode(pts)
    Wp = _w(sd, "decoder.fc_p.weight")[:, :, 0]
    net = e @ Wp.T + _w(sd, "decoder.fc_p.bias")
    masks = []  # (mask, scale) per CBN+ReLU, in forward order
    mats = []
    for i in range(5):
        pre = f"decoder.blocks.{i}"
        a0, s0 = _cbn(sd, pre + ".bn_0", net, lat)
        m0 = a0 > 0
        W0 = _w(sd, pre + ".fc_0.weight")[:, :, 0]
        h = np.maximum(a0, 0) @ W0.T + _w(sd, pre + ".fc_0.bias")
        a1, s1 = _cbn(sd, pre + ".bn_1", h, lat)
        m1 = a1 > 0
        W1 = _w(sd, pre + ".fc_1.weight")[:, :, 0]
        net = net + np.maximum(a1, 0) @ W1.T + _w(sd, pre + ".fc_1.bias")
        masks.append((m0, s0, m1, s1))
        mats.append((W0, W1))
    af, sf = _cbn(sd, "decoder.bn", net, lat)
    mf = af > 0
    wout = _w(sd, "decoder.fc_out.weight")[0, :, 0]
    logit = (np.maximum(af, 0) @ wout + _w(sd, "decoder.fc_out.bias")[0]).astype(f32)
    p = (f32(1.0) / (f32(1.0) + np.exp(-logit, dtype=f32))).astype(f32)
    udf = ((f32(1.0) - p) * f32(0.1)).astype(f32)
    if not want_grad:
        return udf
    # manual reverse mode (what torch.autograd does for udf.sum().backward())
    dlogit = ((f32(-0.1) * (f32(1.0) - p)) * p).astype(f32)       # sigmoid backward
    g = (dlogit[:, None] * wout[None, :]) * mf * sf                  # d/d net (final CBN+ReLU)
    for i in reversed(range(5)):
        m0, s0, m1, s1 = masks[i]
        W0, W1 = mats[i]
        dh = (g @ W1) * m1 * s1
        g = g + (dh @ W0) * m0 * s0
    de = (g @ Wp).astype(f32)
    dx = de[:, 0:3].copy()
    for j in range(10):
        f = f32(2.0 ** j)
        arg = pts * f
        dx += de[:, 3 + 6 * j:6 + 6 * j] * (np.cos(arg, dtype=f32) * f)
        dx += de[:, 6 + 6 * j:9 + 6 * j] * (-np.sin(arg, dtype=f32) * f)
    nrm = np.sqrt((dx * dx).sum(-1, keepdims=True))
    grads = -(dx / np.maximum(nrm, f32(1e-12)))
    return udf, grads.astype(f32)
