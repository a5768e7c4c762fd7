"""diffusion/respace.py of the reference"""
from ...diffusion import SpacedDiffusion, space_timesteps  # noqa: F401
