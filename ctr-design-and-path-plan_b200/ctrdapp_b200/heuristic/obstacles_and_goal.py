from .heuristic import SquareObstacleAvgPlusWeightedGoal  # noqa: F401  (module name of the reference)
