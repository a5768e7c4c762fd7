"""utils/utils.py of the reference: the mesh container handed to the .obj writer (:79-121)"""
from ...output import get_o3d_mesh_from_tensors  # noqa: F401


class GridFiller:
    """utils/utils.py:151-339 -- the udf-only coarse-to-fine filler of the `--watertight` branch (generate_text.py:132-136):
    `GridFiller(size).fill_grid(udf_func, max_batch)` -> (udf [N,N,N], None).  `udf_func` is the scripts' closure over
    surfd_b200's CbnDecoder and one latent (recognised and verified like get_mesh_from_udf does); the lattice is produced by
    surfd_udf_lattice with grad_dev = NULL and stays on the device."""

    def __init__(self, final_resolution, voxel_origin=(-1, -1, -1), cube_side_length=2.0):
        if tuple(float(v) for v in voxel_origin) != (-1.0, -1.0, -1.0) or float(cube_side_length) != 2.0:
            raise NotImplementedError("the lattice kernels hard-code the [-1, 1]^3 cube of the scripts")
        self.N_max = int(final_resolution)

    def fill_grid(self, udf_func, max_batch=2 ** 16):
        from ...meshudf import DecoderUdf, _recognise_closure
        bound = udf_func if isinstance(udf_func, DecoderUdf) else (_recognise_closure(udf_func, 0.1) if callable(udf_func) else None)
        if bound is None:
            raise TypeError("GridFiller.fill_grid needs a DecoderUdf or a udf_func closure over surfd_b200's CbnDecoder and one "
                            "latent; arbitrary Python closures cannot run in the CUDA library and there is no CPU fallback")
        bound.bind()
        udf, _, _ = bound.decoder.lattice(self.N_max, use_fast_grid_filler=True, grads=False)
        return udf, None
