"""Host-side cost bookkeeping of the planner (mirror of ctrdapp.heuristic)."""
from .heuristic import (FollowTheLeader, FollowTheLeaderTranslation, FollowTheLeaderWInsertion, Heuristic,  # noqa: F401
                        OnlyGoalDistance, SquareObstacleAvgPlusWeightedGoal)
from .heuristic_factory import create_heuristic_factory  # noqa: F401
