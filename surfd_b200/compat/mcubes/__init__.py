"""`import mcubes` of the `--watertight` branch (sample/generate_text.py:138-139): marching_cubes(volume, isovalue) ->
(vertices float64 [V,3] in index units, triangles [F,3]) as numpy arrays, computed on the device by surfd_b200.watertight
(classic marching cubes; vertex / face ORDER is not PyMCubes' scan order -- parity unpinned, the package is absent here)."""
import numpy as np
import torch


def marching_cubes(volume, isovalue):
    from ...watertight import marching_cubes as _mc
    if torch.is_tensor(volume) and volume.is_cuda:
        vol = volume
    else:
        if not torch.cuda.is_available():
            raise RuntimeError("surfd_b200 has no CPU path: mcubes.marching_cubes runs on the CUDA device")
        vol = torch.as_tensor(volume).to("cuda")
    v, f = _mc(vol, isovalue)
    return v.cpu().numpy(), f.cpu().numpy().astype(np.uint64)
