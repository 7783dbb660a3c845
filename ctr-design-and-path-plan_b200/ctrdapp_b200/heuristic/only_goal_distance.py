from .heuristic import OnlyGoalDistance  # noqa: F401  (module name of the reference)
