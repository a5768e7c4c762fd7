"""GPU (-m gpu): the conditioning encoders on the device against the golden vectors of the reference's CLIP classes
(tests/golden/make_golden_clip.py; seeded ViT-B/32-shaped checkpoint), and the device result against the host result."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from surfd_b200 import synth
from surfd_b200.clip_encoder import ClipEncoder

pytestmark = pytest.mark.gpu


def test_device_encoders_match_reference_golden():
    g = np.load(os.path.join(GOLDEN, "clip_vitb32.npz"))
    enc = ClipEncoder(synth.synth_clip(77), device="cuda")
    images = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(2024))
    img = enc.encode_image(images)
    txt = enc.encode_text(torch.from_numpy(g["tokens"]))
    assert img.is_cuda and txt.is_cuda and img.dtype == torch.float32
    # fp32 GEMMs on the device (torch's default: no TF32 for matmul); embeddings have |x| ~ 0.8
    assert np.abs(img.cpu().numpy() - g["img_emb"]).max() < 1e-3, np.abs(img.cpu().numpy() - g["img_emb"]).max()
    assert np.abs(txt.cpu().numpy() - g["txt_emb"]).max() < 1e-3, np.abs(txt.cpu().numpy() - g["txt_emb"]).max()
    # as the sampler's context: [B,512] device tensor, finite
    assert tuple(txt.shape) == (5, 512) and torch.isfinite(txt).all() and torch.isfinite(img).all()


def test_watertight_branch_with_the_scripts_imports_swapped(tmp_path):
    """sample/generate_text.py:121-158 statement for statement with `surfd_b200.compat` imports: the udf_func closure, utils.GridFiller,
    `udf[udf < 0] = 0`, mcubes.marching_cubes(udf, 0.01), the mesh export and the pymeshlab component filter."""
    from torch import Tensor
    from surfd_b200.compat.AutoEncoder.models.coordsenc import CoordsEncoder
    from surfd_b200.compat.AutoEncoder.models.cbndec import CbnDecoder
    from surfd_b200.compat.utils.utils import GridFiller
    import surfd_b200.compat.mcubes as mcubes
    from surfd_b200 import output as ml
    from surfd_b200.decoder import UdfDecoder
    latent_size, size = 64, 256
    ckpt = synth.synth_ae_poly(latent_size)
    coords_encoder = CoordsEncoder()
    decoder = CbnDecoder(coords_encoder.out_dim, latent_size, 512, 5)
    decoder.load_state_dict(ckpt["decoder"], strict=True)
    decoder = decoder.cuda()
    decoder.eval()
    lat = 0.3 * torch.randn(1, latent_size, generator=torch.Generator().manual_seed(4)).cuda()
    udf_max_dist = 0.1

    def udf_func(c: Tensor) -> Tensor:
        c = coords_encoder.encode(c.unsqueeze(0))
        p = decoder(c, lat).squeeze(0)
        p = torch.sigmoid(p)
        p = (1 - p) * udf_max_dist
        return p

    fast_grid_filler = GridFiller(size)
    udf, _ = fast_grid_filler.fill_grid(udf_func, max_batch=2 ** 16)
    assert _ is None and tuple(udf.shape) == (size, size, size)
    udf[udf < 0] = 0
    # the same lattice as the library's own call
    ref = UdfDecoder(ckpt["decoder"], latent_size)
    ref.set_latent(lat.reshape(-1))
    want, _, _ = ref.lattice(size, True, grads=False)
    assert torch.equal(udf, want.clamp(min=0))
    vertices, faces = mcubes.marching_cubes(udf.detach().cpu().numpy(), 0.01)
    assert vertices.dtype == np.float64 and vertices.shape[1] == 3 and faces.shape[1] == 3 and faces.shape[0] > 5000
    assert 0 <= vertices.min() and vertices.max() <= size - 1
    mesh_path = str(tmp_path / "a-chair_0.obj")
    ml.write_obj_meshlab(mesh_path, torch.from_numpy(vertices), torch.from_numpy(faces.astype(np.int64)))
    ms = ml.MeshSet()
    ms.set_verbosity(False)
    ms.load_new_mesh(mesh_path)
    ms.meshing_remove_connected_component_by_face_number(mincomponentsize=5000)
    ms.save_current_mesh(mesh_path)
    assert os.path.getsize(mesh_path) > 100000
    with pytest.raises(TypeError):
        fast_grid_filler.fill_grid(lambda c: c[:, 0].abs(), max_batch=2 ** 16)
