from .model import Model  # noqa: F401
from .kinematic import Kinematic, calculate_indices  # noqa: F401
from .model_factory import create_model  # noqa: F401
