"""AutoEncoder/models/cbndec.py of the reference"""
from ....modules import CbnDecoder  # noqa: F401
