"""Test-infrastructure stub for `open3d` (imported by the reference's utils/utils.py)."""
class _Geometry:
    class TriangleMesh:
        pass
geometry = _Geometry
