"""B200-native projection operators for pandekan/tomography_alignment.

Host side mirrors the reference's ``utilities/projection_operators.py`` and ``utilities/geometry.py``;
the arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI of include/tomo_b200.h.
"""
from .geometry import Geometry  # noqa: F401
from .projection_operators import ProjectionMatrix, ProjectionOperator, normalise_poses, pose_table  # noqa: F401

__all__ = ["Geometry", "ProjectionMatrix", "ProjectionOperator", "normalise_poses", "pose_table"]
