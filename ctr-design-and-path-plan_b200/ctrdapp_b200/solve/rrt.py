"""RRT over the batched hot path (mirror of ctrdapp/solve/rrt.py).

The reference draws one random control point, steers from its nearest neighbour, solves the model, checks the
collision and inserts the node -- `iteration_number` times, one Python call chain each (rrt.py:90-131).  Here
`batch_size` candidates are drawn per round against the tree as it stands at the start of the round:

    K x (2T random numbers)                                       rrt.py:146-151
    one ctrb_knn_batch                nearest neighbours          rrt.py:154, dynamic_tree.py:66-89
    step / step_rotation, vectorised                              rrt.py:158-161
    one solve_integrate_batch         indices, g-curves, eta/FTL  rrt.py:110-114, kinematic.py:53-80
    one check_collision_batch         (obstacle_min, goal_dist)   rrt.py:115
    survivors inserted in candidate order                         rrt.py:116-131

With batch_size = 1 (the default) the sequence of operations, and with a seeded generator the tree itself, is
the reference's.  With K > 1 candidates of one round do not see each other as neighbours: the search is
statistically, not node-for-node, the same (the reference's generator is unseeded anyway, rrt.py:5).
A model without `solve_integrate_batch` (any object with the reference's `solve_integrate`) or a collision
checker without `check_collision_batch` is driven through the reference-shaped calls, one candidate at a time.
"""
import random as _random
from math import pi, sqrt

import numpy as np

from .dynamic_tree import DynamicTree
from .solver import Solver
from .step import get_single_tube_value, step_batch, step_rotation_batch


def pack_curves(g_curves):
    """list (tubes) of sequences of 4x4 -> (flat [n,4,4], counts per tube)."""
    counts = [len(t) for t in g_curves]
    parts = [np.asarray(t, np.float64).reshape(-1, 4, 4) for t in g_curves if len(t)]
    flat = np.concatenate(parts, 0) if parts else np.zeros((0, 4, 4))
    return flat, counts


class RRT(Solver):
    def __init__(self, model, heuristic_factory, collision_detector, configuration):
        super().__init__(model, heuristic_factory, collision_detector, configuration)
        self.iter_max = configuration.get("iteration_number")
        self.step_bound = configuration.get("step_bound")
        self.rotation_max = configuration.get("rotation_max")
        self.single_tube_control = configuration.get("single_tube_control")
        self.batch_size = max(1, int(configuration.get("batch_size") or 1))
        seed = configuration.get("seed")
        self.rng = _random.Random(seed) if seed is not None else _random.Random()

        tube_lengths = configuration.get("tube_lengths")
        self.insert_max, covered = [], 0
        for i in range(self.tube_num):  # rrt.py:46-52: what each tube can extend beyond the previous one
            self.insert_max.append(tube_lengths[i] - covered)
            covered = tube_lengths[i]

        ftl = [[np.asarray([0, 0, 0, 100, 0, 0])] for _ in range(self.tube_num)]
        init_heuristic = self.heuristic_factory.create(min_obstacle_distance=0.00001, goal_distance=10e10,
                                                       follow_the_leader=ftl, insertion_fraction=self.tube_num)
        init_g_curves = [[np.eye(4)]] * self.tube_num
        scaling = [1, self.step_bound / self.rotation_max]  # rrt.py:69
        self.tree = DynamicTree(self.tube_num, [0] * self.tube_num, [0] * self.tube_num, init_heuristic, init_g_curves,
                                self.iter_max, scaling, ctx=getattr(model, "ctx", None))
        self.rounds = 0
        self._solve()

    # ------------------------------------------------------------------ main loop
    def _solve(self):
        done = 0
        while done < self.iter_max:
            k = min(self.batch_size, self.iter_max - done)
            self._round(k)
            done += k
            self.rounds += 1

    def _single_iteration(self):
        self._round(1)

    def _round(self, k):
        cands = self.calc_new_random_configs(k)
        result = self._integrate_and_check(cands)
        cands["new_rotation_list"] = np.asarray(cands["new_rotation"]).tolist()  # once per round, not once per candidate
        cands["neighbor_list"] = np.asarray(cands["neighbor"]).tolist()
        for c in range(k):
            self._consider(c, cands, result)

    # ------------------------------------------------------------------ sampling + steering (rrt.py:133-165)
    def calc_new_random_configs(self, k):
        T = self.tube_num
        ctrl = np.empty((k, 2 * T))
        extra = np.empty(k)
        for c in range(k):
            for i in range(T):
                ctrl[c, 2 * i] = self.rng.random() * self.insert_max[i]
                ctrl[c, 2 * i + 1] = self.rng.random() * 2 * pi
            extra[c] = self.rng.random() if self.single_tube_control else 0.0
        neighbor = self.tree.nearest_neighbor_batch(ctrl)
        ins_n = np.array([self.tree.nodes[j].insertion for j in neighbor], np.float64).reshape(k, T)
        rot_n = np.array([self.tree.nodes[j].rotation for j in neighbor], np.float64).reshape(k, T)
        if self.single_tube_control:
            for c in range(k):
                node = self.tree.nodes[neighbor[c]]
                parent = self.tree.nodes[node.parent] if node.parent is not None else node
                ctrl[c], _ = get_single_tube_value(list(ctrl[c]), node.insertion, node.rotation, parent.insertion, 0.8,
                                                   extra[c])
        new_ins = step_batch(ins_n, ctrl[:, 0::2], sqrt(self.step_bound))
        new_rot, d_rot = step_rotation_batch(rot_n, ctrl[:, 1::2], sqrt(self.step_bound) * self.rotation_max / self.step_bound)
        return {"neighbor": neighbor, "new_insert": new_ins, "new_rotation": new_rot, "delta_rotation": d_rot,
                "delta_insert": new_ins - ins_n}

    def calc_new_random_config(self):
        """The reference's single-candidate signature (rrt.py:133-165)."""
        c = self.calc_new_random_configs(1)
        j = int(c["neighbor"][0])
        return (c["delta_rotation"][0].tolist(), c["delta_insert"][0].tolist(), c["new_rotation"][0].tolist(),
                c["new_insert"][0].tolist(), j, self.tree.nodes[j].g_curves)

    # ------------------------------------------------------------------ model + collision for a stack of candidates
    def _want_ftl_arrays(self):
        return getattr(self.heuristic_factory, "ftl_mag_columns", None) is not None and \
            not hasattr(self.heuristic_factory, "create_from_magnitude")

    def _integrate_and_check(self, cands, parents=None, need_g_out=True):
        """-> per-candidate lists: g (list of per-tube arrays), insert_indices, true_insertion, ftl (arrays or None),
        ftl_mag (row or None), obs_min, goal_dist.  `parents` overrides the parent node of each candidate."""
        k = len(cands["new_insert"])
        parents = cands["neighbor"] if parents is None else parents
        T = self.tube_num
        out = {key: [None] * k for key in ("g", "insert_indices", "true_insertion", "ftl", "ftl_mag", "obs", "goal")}
        batched = hasattr(self.model, "solve_integrate_batch") and hasattr(self.cd, "check_collision_batch")
        if not batched:
            for c in range(k):
                g, _, idx, true_ins, ftl = self.model.solve_integrate(
                    list(cands["delta_rotation"][c]), list(cands["delta_insert"][c]), list(cands["new_rotation"][c]),
                    list(cands["new_insert"][c]), self.tree.nodes[parents[c]].g_curves, need_g_out=need_g_out)
                out["g"][c], out["insert_indices"][c], out["true_insertion"][c], out["ftl"][c] = g, idx, true_ins, ftl
                if need_g_out:
                    out["obs"][c], out["goal"][c] = self.cd.check_collision(g, self.tube_rad)
            return out
        packed = [self._packed(p) for p in parents]          # (flat curve, nodes per tube) of every parent
        prev_g = np.concatenate([f for f, _ in packed], 0) if packed else np.zeros((0, 4, 4))
        offs_a = np.zeros(k * T + 1, np.int64)
        if packed:
            np.cumsum(np.asarray([n for _, n in packed], np.int64).ravel(), out=offs_a[1:])
        want_arrays = self._want_ftl_arrays()
        res = self.model.solve_integrate_batch(cands["delta_rotation"], cands["delta_insert"], cands["new_rotation"],
                                               cands["new_insert"], prev_g, offs_a,
                                               need_g_out=need_g_out, want_ftl=want_arrays)
        # whole-array conversions once per round (indexing numpy scalars one by one is what this loop used to spend its time on)
        out["insert_indices"] = res["insert_indices"].tolist()
        out["true_insertion"] = res["true_insertions"].tolist()
        out["ftl_mag"] = list(res["ftl_mag"])
        if need_g_out:
            obs, goal = self.cd.check_collision_batch(res["g"], res["offsets"], self.tube_rad, tube_num=T)
            out["obs"], out["goal"] = obs.tolist(), goal.tolist()
            o, g_all = res["offsets"].tolist(), res["g"]
            out["g"] = [[g_all[o[b + n]:o[b + n + 1]] for n in range(T)] for b in range(0, k * T, T)]
        if want_arrays:
            # kin_eta_ftl_kernel writes npts - prev_idx rows per tube (kinematic.py:181-193); the parent's curve holds
            # exactly that many nodes unless the index rounding moved (then the library reports an error)
            offs, prev, npts = offs_a.tolist(), res["prev_indices"].tolist(), self.model.num_discrete_points
            for c in range(k):
                out["ftl"][c] = [res["ftl"][offs[c * T + n]:offs[c * T + n] + npts[n] - prev[c][n]] for n in range(T)]
        return out

    def _packed(self, node_index):
        node = self.tree.nodes[node_index]
        cached = getattr(node, "_packed", None)
        if cached is None:
            cached = node._packed = pack_curves(node.g_curves)
        return cached

    def _make_heuristic(self, result, c, obs_min, goal_dist, insert_frac):
        f = self.heuristic_factory
        if result["ftl"][c] is None and result["ftl_mag"][c] is not None and hasattr(f, "create_from_magnitude"):
            return f.create_from_magnitude(result["ftl_mag"][c], insertion_fraction=insert_frac)
        return f.create(min_obstacle_distance=obs_min, goal_distance=goal_dist, follow_the_leader=result["ftl"][c],
                        insertion_fraction=insert_frac)

    def _insert_fraction(self, true_insertion):
        # = 0 if all tubes are fully inserted; = tube_num if fully retracted (rrt.py:124-125)
        return sum(1 - (float(i) / ins_max) for i, ins_max in zip(true_insertion, self.insert_max))

    # ------------------------------------------------------------------ what happens to one evaluated candidate
    def _consider(self, c, cands, result):
        obs_min, goal_dist = result["obs"][c], result["goal"][c]
        if obs_min < 0:
            return
        new_index = len(self.tree.nodes)
        if goal_dist < 0:
            self.tree.at_goal.append(new_index)
            self.found_solution = True
            goal_dist = 0
        true_insertion = result["true_insertion"][c]
        heuristic = self._make_heuristic(result, c, obs_min, goal_dist, self._insert_fraction(true_insertion))
        rotation = cands["new_rotation_list"][c] if "new_rotation_list" in cands else cands["new_rotation"][c].tolist()
        parent = cands["neighbor_list"][c] if "neighbor_list" in cands else int(cands["neighbor"][c])
        self.tree.insert(true_insertion, rotation, parent, heuristic, result["g"][c], result["insert_indices"][c])

    # ------------------------------------------------------------------ results (rrt.py:167-200)
    def get_path(self, index):
        g_out = self.tree.get_tube_curves(index)
        insert, rotate, insert_indices = self.tree.get_tube_data(index)
        return g_out, insert, rotate, insert_indices

    def get_best_cost(self):
        pool = self.tree.at_goal if self.found_solution else range(len(self.tree.nodes))
        best_cost, best_index = 1e200, 0
        for i in pool:
            cost = self.tree.nodes[i].get_cost()
            if cost < best_cost:
                best_cost, best_index = cost, i
        return best_cost, best_index

    def save_best_solution(self, output_dir):
        """solution_path.txt in the reference's format (rrt.py:184-198)."""
        cost, best_index = self.get_best_cost()
        chain = list(reversed(self.tree.get_index_list(best_index)))
        with open(output_dir / "solution_path.txt", "w") as f:
            f.write(f"Q: {self.model.q}\n")
            f.write(f"Cost: {cost}\n")
            f.write(", ".join(str(i) for i in chain))

    def save_tree(self, output_dir):
        self.tree.save_tree(output_dir)
