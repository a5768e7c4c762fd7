"""AutoEncoder/models/coordsenc.py of the reference"""
from ....modules import CoordsEncoder  # noqa: F401
