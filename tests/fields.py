"""Analytic UDF fields (udf [N,N,N] float32, grads [N,N,N,3] = -normalize(d udf/dx), zero far from the surface)
used by the marching-cubes parity tests and by tests/golden/make_golden.py.  Pure numpy, seeded."""
import numpy as np


def analytic_field(kind, N, noise=0.0, seed=0):
    rng = np.random.default_rng(seed)
    ax = np.linspace(-1, 1, N).astype(np.float32)
    X, Y, Z = np.meshgrid(ax, ax, ax, indexing="ij")
    P = np.stack([X, Y, Z], -1).astype(np.float64)
    if kind == "sphere":
        r = np.linalg.norm(P, axis=-1); d = r - 0.5; grad = P / np.maximum(r, 1e-9)[..., None]
    elif kind == "torus":
        rho = np.sqrt(P[..., 0] ** 2 + P[..., 1] ** 2)
        q = rho - 0.55; r = np.sqrt(q * q + P[..., 2] ** 2); d = r - 0.2
        gq = np.stack([P[..., 0], P[..., 1], 0 * q], -1) / np.maximum(rho, 1e-9)[..., None]
        grad = (q[..., None] * gq + np.stack([0 * q, 0 * q, P[..., 2]], -1)) / np.maximum(r, 1e-9)[..., None]
    elif kind == "two":
        c = np.array([0.3, 0, 0])
        r1 = np.linalg.norm(P - c, axis=-1); r2 = np.linalg.norm(P + c, axis=-1)
        d1 = r1 - 0.3; d2 = r2 - 0.3
        sel = np.abs(d1) < np.abs(d2)
        d = np.where(sel, d1, d2)
        grad = np.where(sel[..., None], (P - c) / np.maximum(r1, 1e-9)[..., None], (P + c) / np.maximum(r2, 1e-9)[..., None])
    elif kind == "hemi":
        # open surface: the cap x > 0.1 of the sphere r = 0.5 (a surface with boundary)
        r = np.linalg.norm(P, axis=-1); d_s = np.abs(r - 0.5)
        rc = np.sqrt(0.25 - 0.01)
        rho = np.sqrt(P[..., 1] ** 2 + P[..., 2] ** 2)
        d_c = np.sqrt((P[..., 0] - 0.1) ** 2 + (rho - rc) ** 2)
        oncap = P[..., 0] / np.maximum(r, 1e-9) * 0.5 > 0.1
        udf = np.where(oncap, d_s, d_c)
        a64 = ax.astype(np.float64)
        gx, gy, gz = np.gradient(udf, a64, a64, a64)
        grad = np.stack([gx, gy, gz], -1)
        grad = grad / np.maximum(np.linalg.norm(grad, axis=-1, keepdims=True), 1e-12)
        d = None
    else:
        raise ValueError(kind)
    if d is not None:
        udf = np.abs(d); grad = grad * np.sign(d)[..., None]
    udf = udf.astype(np.float32)
    g = (-grad).astype(np.float32)
    if noise > 0:
        g = g + noise * rng.standard_normal(g.shape).astype(np.float32)
        g = (g / np.maximum(np.linalg.norm(g, axis=-1, keepdims=True), 1e-12)).astype(np.float32)
        udf = np.abs(udf + (0.2 * noise * 2 / N) * rng.standard_normal(udf.shape).astype(np.float32)).astype(np.float32)
    g[udf > 2.5 * 2 / N] = 0
    return np.ascontiguousarray(udf), np.ascontiguousarray(g)


MC_CASES = [("sphere", 32, 0.0), ("sphere", 33, 0.3), ("torus", 48, 0.3), ("two", 40, 0.0), ("hemi", 40, 0.0),
            ("hemi", 64, 1.0), ("torus", 64, 1.0)]


def noise_field(N, seed, level=0.3):
    """every cube is a candidate (udf = level * voxel everywhere) and the gradients are random unit vectors: the walk is all
    seeds, unsure pushes and ambiguous cases -- the stress case for queues, priorities and capacities"""
    rng = np.random.default_rng(seed)
    udf = np.full((N, N, N), level * 2.0 / (N - 1), np.float32)
    udf *= rng.uniform(0.5, 1.5, udf.shape).astype(np.float32)
    g = rng.standard_normal((N, N, N, 3)).astype(np.float32)
    g /= np.linalg.norm(g, axis=-1, keepdims=True)
    return udf, np.ascontiguousarray(g)
