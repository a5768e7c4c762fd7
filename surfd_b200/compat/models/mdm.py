"""models/mdm.py of the reference"""
from ...diffusion import MDM  # noqa: F401
