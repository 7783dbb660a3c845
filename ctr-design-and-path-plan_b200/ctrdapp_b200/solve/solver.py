"""Planner interface (mirror of ctrdapp/solve/solver.py:4-202 without the plotting hooks, which need pyvista)."""
from abc import ABC, abstractmethod


class Solver(ABC):
    def __init__(self, model, heuristic_factory, collision_detector, configuration):
        self.model = model
        self.tube_num = configuration.get("tube_number")
        self.tube_rad = configuration.get("tube_radius").get("outer")
        self.heuristic_factory = heuristic_factory
        self.cd = collision_detector
        self.found_solution = False

    @abstractmethod
    def get_path(self, index):
        ...

    @abstractmethod
    def get_best_cost(self):
        ...

    @abstractmethod
    def save_best_solution(self, output_dir):
        ...

    @abstractmethod
    def save_tree(self, output_dir):
        ...
