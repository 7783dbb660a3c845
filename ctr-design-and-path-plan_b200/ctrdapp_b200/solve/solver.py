"""What a planner object offers the optimiser and the result writers (method names of ctrdapp/solve/solver.py:4-202;
the pyvista plotting hooks of the reference are out of scope)."""


class Solver:
    found_solution = False

    def __init__(self, model, heuristic_factory, collision_detector, configuration):
        self.model, self.heuristic_factory, self.cd = model, heuristic_factory, collision_detector
        self.tube_num = configuration.get("tube_number")
        self.tube_rad = configuration.get("tube_radius").get("outer")
        self.found_solution = False

    def _missing(self, what):
        raise NotImplementedError(f"{type(self).__name__}.{what}")

    def get_path(self, index):
        """(g-curves, insertions, rotations[, insert indices]) from node `index` back to the root"""
        self._missing("get_path")

    def get_best_cost(self):
        """(cost, node index) of the best node, nodes at the goal first"""
        self._missing("get_best_cost")

    def save_best_solution(self, output_dir):
        self._missing("save_best_solution")

    def save_tree(self, output_dir):
        self._missing("save_tree")
