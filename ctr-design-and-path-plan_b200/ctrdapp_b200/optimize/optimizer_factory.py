"""create_optimizer (mirror of ctrdapp/optimize/optimizer_factory.py:10-37)."""
from .neldermead import NelderMead, Optimizer


def create_optimizer(heuristic_factory, collision_checker, initial_guess, configuration, **kwargs) -> Optimizer:
    name = configuration.get("optimizer_type")
    if name == "nelder_mead":
        return NelderMead(heuristic_factory, collision_checker, initial_guess, configuration, **kwargs)
    raise UserWarning(f"{name} is not a defined optimizer. Change config file.")
