"""Test-infrastructure stub for trimesh (reference meshudf/meshudf.py imports it at top)."""
