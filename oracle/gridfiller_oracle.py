"""ORACLE (test infrastructure; never imported by the product path).

numpy restatement of the reference's lattice samplers, for any `udf_func(points[M,3]) -> udf[M]` and
`grad_func(points) -> grads[M,3]`:
  GridFiller.__init__ / fill_grid     meshudf/meshudf.py:36-121, 123-206   (mask formulation kept literally)
  get_udf_and_grads                   meshudf/meshudf.py:254-304
Coordinates are built as the reference does: float32(idx) * float32(voxel) then + (-1) (two roundings).
Pinned against the reference's own GridFiller in tests/golden/make_golden.py (gridfiller_*.npz).
"""
import math
import numpy as np

f32 = np.float32


def lattice_coords(N):
    idx = np.arange(N ** 3, dtype=np.int64)
    voxel = f32(2.0 / (N - 1))
    s = np.zeros((N ** 3, 3), f32)
    s[:, 2] = (idx % N).astype(f32)
    s[:, 1] = ((idx // N) % N).astype(f32)
    s[:, 0] = ((idx // N // N) % N).astype(f32)
    s = (s * voxel).astype(f32)
    return (s + f32(-1.0)).astype(f32)


def fill_grid(udf_func, grad_func, N):
    """returns udf [N,N,N], grads [N,N,N,3], dict(n_udf, n_grad, per_level)"""
    levels = [32 * (2 ** i) for i in range(int(math.log2(N) - 4))]
    coords = lattice_coords(N)
    udf = np.zeros(N ** 3, f32)
    mask0 = np.zeros((N, N, N), bool)
    masks_coarse, blocks, no_recompute = {}, {}, {}
    for i, NL in enumerate(levels):
        S = N // NL
        mc = mask0.copy(); mc[::S, ::S, ::S] = True
        mc = mc.reshape(-1)
        masks_coarse[i] = mc
        nb = mask0.copy(); nb[:S, :S, :S] = True
        blocks[i] = np.where(mc)[0].reshape(-1, 1) + np.where(nb.reshape(-1))[0].reshape(1, -1)
        if i > 0:
            nr = mc.copy(); nr[masks_coarse[i - 1]] = False
            no_recompute[i] = nr
    close, local_blocks, per_level = {}, {}, []
    for level, NL in enumerate(levels):
        if level == 0:
            mask_coarse = masks_coarse[0]
            blk = blocks[0]
            mask_nr = masks_coarse[0]
        else:
            mask_coarse = masks_coarse[level].copy()
            for l in range(level):
                mask_coarse[local_blocks[l][~close[l]].reshape(-1)] = False
            blk = blocks[level][mask_coarse[masks_coarse[level]]] if NL < N else blocks[level]
            mask_nr = no_recompute[level].copy()
            for l in range(level):
                mask_nr[local_blocks[l][~close[l]].reshape(-1)] = False
        local_blocks[level] = blk
        udf[mask_nr] = udf_func(coords[mask_nr])
        per_level.append(int(mask_nr.sum()))
        if NL < N:
            thr = f32(1.5 * 1.7 * (2.0 / NL))
            cl = np.abs(udf[mask_coarse]) < thr
            close[level] = cl
            far_blocks = blk[~cl]
            udf[far_blocks] = udf[mask_coarse][~cl][:, None]
    grads = np.zeros((N ** 3, 3), f32)
    mg = udf < f32(2.5 * 2.0 / N)
    if mg.any():
        grads[mg] = grad_func(coords[mg])
    return udf.reshape(N, N, N), grads.reshape(N, N, N, 3), dict(n_udf=sum(per_level), n_grad=int(mg.sum()), per_level=per_level)


def dense_grid(udf_func, grad_func, N, max_dist=0.1):
    coords = lattice_coords(N)
    udf = udf_func(coords).astype(f32)
    grads = np.zeros((N ** 3, 3), f32)
    mg = udf < f32(max_dist - 1e-3)
    if mg.any():
        grads[mg] = grad_func(coords[mg])
    return udf.reshape(N, N, N), grads.reshape(N, N, N, 3), dict(n_udf=N ** 3, n_grad=int(mg.sum()))
