"""utils/model_util.py of the reference"""
from ...diffusion import (create_gaussian_diffusion, create_model_and_diffusion, get_model_args,  # noqa: F401
                          load_model_wo_clip)
