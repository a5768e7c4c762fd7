"""models/cfg_sampler.py of the reference"""
from ...diffusion import ClassifierFreeSampleModel  # noqa: F401
