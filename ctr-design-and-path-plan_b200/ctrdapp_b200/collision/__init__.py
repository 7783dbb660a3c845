"""Mirror of ctrdapp/collision/__init__.py."""
from .collision_checker import CollisionChecker  # noqa: F401
from .init_collision import add_goal, add_obstacles, parse_json, parse_mesh, rotate_align  # noqa: F401
