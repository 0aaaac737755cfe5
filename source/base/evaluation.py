"""source.base.evaluation -> points2surf_b200.evaluation (eval_predictions, mesh_comparison, ...)."""
from points2surf_b200.evaluation import *  # noqa: F401,F403
from points2surf_b200.evaluation import (eval_predictions, mesh_comparison, compare_predictions_binary_tensors,  # noqa: F401
                                         print_list_of_dicts)
