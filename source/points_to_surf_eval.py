"""source.points_to_surf_eval -> points2surf_b200.eval (parse_arguments, points_to_surf_eval)."""
from points2surf_b200.eval import *  # noqa: F401,F403
from points2surf_b200.eval import parse_arguments, points_to_surf_eval  # noqa: F401
