"""Design optimisation (mirror of ctrdapp.optimize) with vertex-parallel Nelder-Mead."""
from .neldermead import NelderMead, calculate_simplex  # noqa: F401
from .optimizer_factory import create_optimizer  # noqa: F401
from .simplex import DistributedEvaluator, nelder_mead  # noqa: F401
