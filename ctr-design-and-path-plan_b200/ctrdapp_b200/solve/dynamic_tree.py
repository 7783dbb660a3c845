"""Node store of the planner (mirror of ctrdapp/solve/dynamic_tree.py).

The reference keeps the nodes in a Python list and their [insertion, rotation, ...] points in pympnn's k-d
trees (dynamic_tree.py:45-61).  Here the points live in one growing array that is mirrored on the device, and
the neighbour queries are a batched brute-force scan (`ctrb_knn_batch`, csrc/ctrb_knn.cu): the tree never holds
more than iteration_number + 2 points, and the planner asks for hundreds of neighbours at a time.

Metric (parity unpinned: pympnn / MPNN is a third-party C++ library that is not in the tree and cannot be
installed offline).  What the reference fixes: topology [1, 2] * T (dynamic_tree.py:53), `scaling` * T as the
per-dimension scale, distances compared against `max_distance` as returned.  The only literals are
test/test_solve.py:83-92 (scaling [1, 6]); they hold for the Euclidean distance in which a circular difference
is DIVIDED by its scale -- rotation barely matters, as the docstring of RRT._single_iteration says ("the
nearest neighbor is only found by insertion values") -- and for no reading in which it is multiplied.  That is
the default here; `rotation_metric="multiply"` selects the other reading.
"""
import numpy as np

from .. import _lib as L


class Node:
    """dynamic_tree.py:358-397."""

    def __init__(self, insertion, rotation, dim, heuristic, g_curves, insert_indices, parent=None):
        if len(insertion) != dim or len(rotation) != dim:
            raise ValueError(f"Data size does not match given tube number ({dim})")
        self.insertion = insertion
        self.rotation = rotation
        self.parent = parent
        self.heuristic = heuristic
        self.g_curves = g_curves
        self.insert_indices = insert_indices
        self.children = []

    def get_cost(self):
        return self.heuristic.get_cost()


def _interleave(ins, rot):
    pt = np.empty(2 * len(ins))
    pt[0::2] = ins
    pt[1::2] = rot
    return pt


class DynamicTree:
    def __init__(self, dim, ins, rot, init_heuristic, init_g_curves, max_tree_size, scaling, ctx=None,
                 rotation_metric="divide"):
        self.tube_num = dim
        self.ctx = ctx or L.default_context()
        init_insert_indices = [len(init_g_curves[0]) - 1] * self.tube_num
        self.nodes = [Node(ins, rot, dim, init_heuristic, init_g_curves, init_insert_indices)]
        self.at_goal = []
        self.topology = np.ascontiguousarray([1, 2] * dim, np.int32)
        s_lin, s_rot = float(scaling[0]), float(scaling[1])
        w_rot = 1.0 / s_rot if rotation_metric == "divide" else s_rot
        self.weight = np.ascontiguousarray([s_lin, w_rot] * dim, np.float64)
        self._pts = np.empty((max(int(max_tree_size) + 2, 16), 2 * dim))
        self._n = 0
        self._append_point(ins, rot)

    # ------------------------------------------------------------------ points
    def _append_point(self, ins, rot):
        if self._n == len(self._pts):
            self._pts = np.concatenate([self._pts, np.empty_like(self._pts)], 0)
        self._pts[self._n] = _interleave(ins, rot)
        self._n += 1

    def knn(self, queries, k):
        """k nearest stored points of each query: ([Q, k] indices (-1 = none), [Q, k] distances)."""
        q = np.ascontiguousarray(np.asarray(queries, np.float64).reshape(-1, 2 * self.tube_num))
        idx = np.empty((len(q), k), np.int32)
        dist = np.empty((len(q), k))
        pts = np.ascontiguousarray(self._pts[:self._n])
        L.check(L.lib().ctrb_knn_batch(self.ctx.handle, 2 * self.tube_num, self._n, L.ptr(pts), len(q), L.ptr(q),
                                       L.ptr(self.weight), L.ptr(self.topology), k, L.ptr(idx), L.ptr(dist),
                                       L.MEM_HOST))
        return idx, dist

    # ------------------------------------------------------------------ queries (dynamic_tree.py:66-120)
    def _neighbor_tuple(self, node_ind):
        node = self.nodes[node_ind]
        parent = self.nodes[node.parent] if node.parent is not None else node
        return node.insertion, node.rotation, node_ind, parent.insertion

    def nearest_neighbor(self, x):
        idx, _ = self.knn([x], 1)
        return self._neighbor_tuple(int(idx[0, 0]))

    def nearest_neighbor_batch(self, X):
        """Nearest node index of every row of X against the tree as it stands (one launch)."""
        idx, _ = self.knn(X, 1)
        return idx[:, 0].astype(np.int64)

    @staticmethod
    def _within(idx_row, dist_row, max_distance):
        """dynamic_tree.py:107-120: the prefix of the sorted neighbours within max_distance, never empty."""
        keep = 0
        for d, i in zip(dist_row, idx_row):
            if i < 0 or d > max_distance:
                break
            keep += 1
        return [int(i) for i in idx_row[:max(keep, 1)]]

    def find_all_nearest_neighbor(self, x, max_distance):
        return self.find_all_nearest_neighbor_batch([x], max_distance)[0]

    def find_all_nearest_neighbor_batch(self, X, max_distance):
        """The reference asks for 4, then 8, then 16 neighbours while the farthest is still inside
        (dynamic_tree.py:101-105); one k = 16 query returns the same prefix."""
        k = min(16, max(self._n, 1))
        idx, dist = self.knn(X, k)
        out = []
        for r in range(len(idx)):
            n_in = int(np.sum((idx[r] >= 0) & (dist[r] <= max_distance)))
            kk = 4
            while kk < 16 and n_in >= kk:  # the while loop of :101-105 stops at the first k whose farthest is outside
                kk *= 2
            out.append(self._within(idx[r, :kk], dist[r, :kk], max_distance))
        return out

    # ------------------------------------------------------------------ structure (dynamic_tree.py:122-262)
    def insert(self, ins, rot, parent, heuristic, g_curves, insert_indices):
        init_insertion = set(ins) == {0.0}
        heuristic.calculate_cost_from_parent(self.nodes[parent].heuristic, init_insertion=init_insertion)
        self.nodes.append(Node(ins, rot, self.tube_num, heuristic, g_curves, insert_indices, parent))
        self.nodes[parent].children.append(len(self.nodes) - 1)
        self._append_point(ins, rot)

    def no_cycle(self, parent_ind, child_ind):
        """True when child_ind is not an ancestor of parent_ind (iterative; the reference recurses)."""
        if self.nodes[parent_ind].parent is None:
            raise ValueError('Parent_ind should never be entered as 0 (root node).')
        ancestor = self.nodes[parent_ind].parent
        while True:
            if ancestor == child_ind:
                return False
            if ancestor == 0:
                return True
            ancestor = self.nodes[ancestor].parent

    def reset_heuristic_all_children(self, ind):
        stack = [ind]
        while stack:
            p = stack.pop()
            ph = self.nodes[p].heuristic
            for ch in self.nodes[p].children:
                node = self.nodes[ch]
                node.heuristic.calculate_cost_from_parent(ph, reset=True, init_insertion=set(node.insertion) == {0.0})
                stack.append(ch)

    def swap_parents(self, current_ind, new_parent_ind, new_heuristic):
        node = self.nodes[current_ind]
        self.nodes[node.parent].children.remove(current_ind)
        node.heuristic = new_heuristic
        node.parent = new_parent_ind
        self.nodes[new_parent_ind].children.append(current_ind)
        new_heuristic.calculate_cost_from_parent(self.nodes[new_parent_ind].heuristic, reset=True,
                                                 init_insertion=set(node.insertion) == {0.0})
        self.reset_heuristic_all_children(current_ind)

    # ------------------------------------------------------------------ read-outs (dynamic_tree.py:264-355)
    def get_index_list(self, child_ind):
        out = [child_ind]
        while out[-1] != 0:
            out.append(self.nodes[out[-1]].parent)
        return out

    def get_costs(self, child_ind):
        return [self.nodes[i].get_cost() for i in self.get_index_list(child_ind)]

    def get_tube_data(self, child_ind):
        chain = self.get_index_list(child_ind)
        return ([self.nodes[i].insertion for i in chain], [self.nodes[i].rotation for i in chain],
                [self.nodes[i].insert_indices for i in chain])

    def get_tube_curves(self, child_ind):
        return [self.nodes[i].g_curves for i in self.get_index_list(child_ind)]

    def save_tree(self, output_dir):
        """tree.txt in the reference's format: "index | insertions | rotations | parent | cost | at_goal"."""
        goal = set(self.at_goal)
        with open(output_dir / "tree.txt", "w") as f:
            for i, node in enumerate(self.nodes):
                ins = ", ".join(str(v) for v in node.insertion)
                rot = ", ".join(str(v) for v in node.rotation)
                f.write(f"{i} | {ins} | {rot} | {node.parent} | {node.heuristic.get_cost()} | {i in goal}\n")
