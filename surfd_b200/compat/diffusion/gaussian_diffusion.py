"""diffusion/gaussian_diffusion.py of the reference (the part the sampling scripts reach)"""
from ...diffusion import get_named_beta_schedule  # noqa: F401
