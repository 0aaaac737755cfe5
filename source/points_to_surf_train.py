"""source.points_to_surf_train -> points2surf_b200.points_to_surf_train (parse_arguments, points_to_surf_train)."""
from points2surf_b200.points_to_surf_train import *  # noqa: F401,F403
from points2surf_b200.points_to_surf_train import parse_arguments, points_to_surf_train  # noqa: F401
