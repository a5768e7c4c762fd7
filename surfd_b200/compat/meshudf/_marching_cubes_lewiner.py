"""meshudf/_marching_cubes_lewiner.py of the reference"""
from ...meshudf import udf_mc_lewiner  # noqa: F401
