"""Module paths of the reference, re-exporting the B200 drop-ins: a script of the reference runs with its imports swapped
from `X` to `surfd_b200.compat.X` (e.g. `from surfd_b200.compat.utils.model_util import create_model_and_diffusion`).
INTEGRATION.md lists the mapping; tests/test_compat_flow.py runs the generate_uncond flow through it."""
