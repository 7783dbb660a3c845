"""create_solver (mirror of ctrdapp/solve/solver_factory.py:9-37)."""
from .rrt import RRT
from .rrt_star import RRTStar
from .solver import Solver
from .tree_from_file import TreeFromFile


def create_solver(model, heuristic_factory, collision_detector, configuration) -> Solver:
    name = configuration.get("solver_type")
    table = {"rrt": RRT, "rrt_star": RRTStar, "tree_from_file": TreeFromFile}
    if name not in table:
        raise UserWarning(f"{name} is not a defined solver. Change config file.")
    return table[name](model, heuristic_factory, collision_detector, configuration)
