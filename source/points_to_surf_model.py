"""source.points_to_surf_model -> points2surf_b200.model (PointsToSurfModel: same constructor, same state_dict keys)."""
from points2surf_b200.model import PointsToSurfModel  # noqa: F401
