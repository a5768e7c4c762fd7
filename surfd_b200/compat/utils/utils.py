"""utils/utils.py of the reference: the mesh container handed to the .obj writer (:79-121)"""
from ...output import get_o3d_mesh_from_tensors  # noqa: F401
