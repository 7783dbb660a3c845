"""The planner that calls the hot path (mirror of ctrdapp.solve): batched RRT / RRT*, the node store with
device-side nearest neighbours, steering arithmetic."""
from .dynamic_tree import DynamicTree, Node  # noqa: F401
from .rrt import RRT  # noqa: F401
from .rrt_star import RRTStar  # noqa: F401
from .solver_factory import create_solver  # noqa: F401
