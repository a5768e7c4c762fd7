"""meshudf/meshudf.py of the reference"""
from ...meshudf import get_mesh_from_udf  # noqa: F401
