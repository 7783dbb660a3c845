"""Rebuild a tree from a saved tree.txt (mirror of ctrdapp/solve/tree_from_file.py): every stored node's g-curves
are re-solved -- here as ONE solve_g_batch over all nodes instead of one solve_g per node."""
import pathlib
from collections import deque

from ..heuristic.heuristic_factory import create_heuristic_factory
from .rrt import RRT


def _to_bool(text):
    t = text.strip().lower()
    if t in ("y", "yes", "t", "true", "on", "1"):
        return True
    if t in ("n", "no", "f", "false", "off", "0"):
        return False
    raise ValueError(f"invalid truth value {text!r}")


def _order_nodes(file):
    """Nodes of tree.txt in an order in which every parent precedes its children (breadth first), with parent
    indices rewritten to positions in that order (tree_from_file.py:50-106)."""
    children, nodes = {}, []
    with open(file, "r") as tree:
        for line in tree:
            f = line.split(" | ")
            ind = int(f[0])
            parent = None if ind == 0 else int(f[3])
            # the reference parses insertions with int() (tree_from_file.py:28), which only reads back trees whose tube
            # lengths and delta_x are integers; fractional insertions ("2.5") are accepted here as well
            ins = [float(v) for v in f[1].split(", ")]
            ins = [int(v) if v == int(v) else v for v in ins]
            nodes.append({"ins": ins, "rot": [float(v) for v in f[2].split(", ")],
                          "parent": parent, "cost": float(f[4]), "goal": _to_bool(f[5])})
            children.setdefault(parent, []).append(ind)
    order = deque(children.get(0, []))
    queue = deque(order)
    position = {ind: k + 1 for k, ind in enumerate(order)}
    while queue:
        ind = queue.popleft()
        for ch in children.get(ind, []):
            position[ch] = len(order) + 1
            order.append(ch)
            queue.append(ch)
            nodes[ch]["parent"] = position[ind]
    return order, nodes


class TreeFromFile(RRT):
    def __init__(self, model, heuristic_factory, collision_detector, configuration):
        filename = pathlib.Path().absolute() / "output" / configuration.get("run_identifier") / "tree.txt"
        self.node_order, self.nodes_list = _order_nodes(filename)
        configuration = dict(configuration)
        configuration["heuristic_type"] = "only_goal_distance"  # costs are taken from the file as they are
        heuristic_factory = create_heuristic_factory(configuration, {"only_goal_distance": []})
        # the reference reads both from the configuration (tree_from_file.py:20-21); the model holds the same values
        self.delta_x = configuration.get("delta_x") or model.delta_x
        self.tube_lengths = configuration.get("tube_lengths") or model.max_tube_length
        super().__init__(model, heuristic_factory, collision_detector, configuration)

    def _round(self, k):
        todo = [self.node_order.popleft() for _ in range(min(k, len(self.node_order)))]
        if not todo:
            return
        nodes = [self.nodes_list[i] for i in todo]
        # insertion indices from the values (inverse insertion, tree_from_file.py:42)
        idx = [[int(round((length - ins) / self.delta_x)) for ins, length in zip(n["ins"], self.tube_lengths)] for n in nodes]
        T = self.tube_num
        if hasattr(self.model, "solve_g_batch"):
            g, off = self.model.solve_g_batch(idx, [n["rot"] for n in nodes], full=False)
            curves = [[g[int(off[c * T + t]):int(off[c * T + t + 1])] for t in range(T)] for c in range(len(nodes))]
        else:
            curves = [self.model.solve_g(i, n["rot"], full=False) for i, n in zip(idx, nodes)]
        for n, i, g_c in zip(nodes, idx, curves):
            if n["goal"]:
                self.tree.at_goal.append(len(self.tree.nodes))
                self.found_solution = True
            self.tree.insert(n["ins"], n["rot"], n["parent"], self.heuristic_factory.create(goal_distance=n["cost"]), g_c, i)
