"""RRT* over the batched hot path (mirror of ctrdapp/solve/rrt_star.py).

Per accepted candidate the reference (rrt_star.py:13-93) (1) collects up to 16 neighbours within step_bound,
(2) re-evaluates the follow-the-leader twist with each of them as the parent -- one solve_integrate call each --
and hangs the node under the cheapest, (3) offers the new node as a parent to the remaining neighbours, again one
solve_integrate each.  Here the expensive part of a round (g-curves + collision of `batch_size` candidates) is one
batch as in RRT, and steps (2) and (3) are one eta/FTL batch each over all (candidate, neighbour) pairs of the
node -- or no model call at all when the cost does not depend on the parent (`heuristic_factory.needs_parent`
False: obstacle/goal costs only need the parent's running statistic).

With `batch_size` > 1 a whole round works against the tree as it stood when the round began (as batched RRT does):
one k-NN launch finds the neighbours of every surviving candidate, ONE eta/FTL batch re-evaluates all (candidate,
alternative parent) pairs, the survivors are inserted, and ONE more batch evaluates all (new node, neighbour) rewiring
offers, which are then applied in candidate order (cycle check and cost comparison at that moment, so a neighbour that an
earlier candidate of the round has already re-parented is judged by its new cost).  New nodes of the same round are not
each other's neighbours.  `batch_size` = 1 follows the reference's sequence of operations, with two deliberate deviations
(so trees are not asserted node for node against the reference's under a seeded generator):
  * the reference sets `single_tube_control = False` only AFTER `super().__init__` has already run the search
    (rrt_star.py:9-11), so the configured value is in effect during its search; here it is off during the search, which is
    what its comment ("impossible to maintain with the rewire operation") asks for;
  * the reference's choose-parent loop reuses the previous neighbour's FTL array for the original neighbour when that
    neighbour is not first in the list (rrt_star.py:43-61); here the original solve's FTL is always used.
"""
import numpy as np

from .rrt import RRT
from .step import delta_rotation_batch


class RRTStar(RRT):
    def __init__(self, model, heuristic_factory, collision_detector, configuration):
        configuration = dict(configuration)
        configuration["single_tube_control"] = False  # impossible to maintain with the rewire operation (rrt_star.py:11; see above)
        super().__init__(model, heuristic_factory, collision_detector, configuration)
        self.single_tube_control = False

    def _ftl_from_parents(self, parents, to_ins, to_rot):
        """FTL of moving from each node in `parents` to the configuration (to_ins[i], to_rot[i]) -> result dict of
        _integrate_and_check without g-curves (need_g_out=False, rrt_star.py:49-51,83-85)."""
        k = len(parents)
        ins_p = np.array([self.tree.nodes[j].insertion for j in parents], np.float64).reshape(k, self.tube_num)
        rot_p = np.array([self.tree.nodes[j].rotation for j in parents], np.float64).reshape(k, self.tube_num)
        cands = {"new_insert": np.asarray(to_ins, np.float64).reshape(k, -1),
                 "new_rotation": np.asarray(to_rot, np.float64).reshape(k, -1)}
        cands["delta_insert"] = cands["new_insert"] - ins_p
        cands["delta_rotation"] = delta_rotation_batch(rot_p, cands["new_rotation"])
        return self._integrate_and_check(cands, parents=list(parents), need_g_out=False)

    def _round(self, k):
        if k == 1:
            return super()._round(1)  # sequential: the reference's order of operations (_consider below)
        tree, f = self.tree, self.heuristic_factory
        cands = self.calc_new_random_configs(k)
        result = self._integrate_and_check(cands)
        parent_dependent = getattr(f, "needs_parent", True)
        live = [c for c in range(k) if result["obs"][c] >= 0]
        if not live:
            return
        # ---- neighbours of every survivor in the tree as it stands: one launch
        T = self.tube_num
        q = np.empty((len(live), 2 * T))
        for s_, c in enumerate(live):
            q[s_, 0::2] = result["true_insertion"][c]
            q[s_, 1::2] = cands["new_rotation"][c]
        neigh = tree.find_all_nearest_neighbor_batch(q, self.step_bound)
        # ---- choose the parents (rrt_star.py:40-67): one eta / FTL batch over all (survivor, alternative parent) pairs
        pairs = []
        if parent_dependent:
            for s_, c in enumerate(live):
                original = int(cands["neighbor"][c])
                pairs += [(s_, j) for j in neigh[s_] if j != original]
        alt = None
        if pairs:
            alt = self._ftl_from_parents([j for _, j in pairs], [result["true_insertion"][live[s_]] for s_, _ in pairs],
                                         [cands["new_rotation"][live[s_]].tolist() for s_, _ in pairs])
        alt_row = {sj: n for n, sj in enumerate(pairs)}
        new_nodes = []  # (new index, its heuristic, its neighbours, chosen parent)
        for s_, c in enumerate(live):
            obs_min, goal_dist = result["obs"][c], result["goal"][c]
            new_index = len(tree.nodes)
            if goal_dist < 0:
                tree.at_goal.append(new_index)
                self.found_solution = True
                goal_dist = 0
            true_insertion = result["true_insertion"][c]
            insert_frac = self._insert_fraction(true_insertion)
            original = int(cands["neighbor"][c])
            best_cost, best_parent, best_heuristic = None, None, None
            for j in neigh[s_]:
                n = alt_row.get((s_, j))
                h = self._make_heuristic(alt, n, obs_min, goal_dist, insert_frac) if n is not None else \
                    self._make_heuristic(result, c, obs_min, goal_dist, insert_frac)
                cost = h.test_cost_from_parent(tree.nodes[j].heuristic)
                if best_cost is None or cost < best_cost:
                    best_cost, best_parent, best_heuristic = cost, j, h
            tree.insert(true_insertion, cands["new_rotation"][c].tolist(), best_parent, best_heuristic, result["g"][c],
                        result["insert_indices"][c])
            new_nodes.append((new_index, best_heuristic, neigh[s_], best_parent))
        # ---- rewiring offers (rrt_star.py:69-93): one batch over all (new node, neighbour) pairs, applied in order
        offers = [(n_i, j) for n_i, (new_index, _, nb, bp) in enumerate(new_nodes) for j in nb if j != bp]
        if not offers:
            return
        re = None
        if parent_dependent:
            re = self._ftl_from_parents([new_nodes[n_i][0] for n_i, _ in offers], [tree.nodes[j].insertion for _, j in offers],
                                        [tree.nodes[j].rotation for _, j in offers])
        for n, (n_i, j) in enumerate(offers):
            new_index, best_heuristic = new_nodes[n_i][0], new_nodes[n_i][1]
            if tree.nodes[j].parent == new_index or not tree.no_cycle(new_index, j):
                continue
            current = tree.nodes[j].heuristic
            if re is not None and re["ftl"][n] is None and hasattr(f, "create_from_magnitude"):
                candidate = f.create_from_magnitude(re["ftl_mag"][n], insertion_fraction=getattr(current, "insertion_fraction", None))
            else:
                candidate = f.create_from_old(current, follow_the_leader=re["ftl"][n] if re is not None else None)
            if candidate.test_cost_from_parent(best_heuristic) < current.get_cost():
                tree.swap_parents(j, new_index, candidate)

    def _consider(self, c, cands, result):
        obs_min, goal_dist = result["obs"][c], result["goal"][c]
        if obs_min < 0:
            return
        tree, f = self.tree, self.heuristic_factory
        new_index = len(tree.nodes)
        if goal_dist < 0:
            tree.at_goal.append(new_index)
            self.found_solution = True
            goal_dist = 0
        true_insertion = result["true_insertion"][c]
        new_rotation = cands["new_rotation"][c].tolist()
        insert_frac = self._insert_fraction(true_insertion)
        original = int(cands["neighbor"][c])
        parent_dependent = getattr(f, "needs_parent", True)

        query_pt = np.empty(2 * self.tube_num)
        query_pt[0::2] = true_insertion
        query_pt[1::2] = new_rotation
        neighbor_indices = tree.find_all_nearest_neighbor(query_pt, self.step_bound)

        # ---- choose the parent (rrt_star.py:40-67)
        others = [j for j in neighbor_indices if j != original] if parent_dependent else []
        alt = self._ftl_from_parents(others, [true_insertion] * len(others), [new_rotation] * len(others)) if others else None
        heur_arr, cost_arr = [], []
        for j in neighbor_indices:
            if alt is not None and j != original:
                h = self._make_heuristic(alt, others.index(j), obs_min, goal_dist, insert_frac)
            else:
                h = self._make_heuristic(result, c, obs_min, goal_dist, insert_frac)
            heur_arr.append(h)
            cost_arr.append(h.test_cost_from_parent(tree.nodes[j].heuristic))
        index_min = min(range(len(cost_arr)), key=cost_arr.__getitem__)
        best_parent, best_heuristic = neighbor_indices[index_min], heur_arr[index_min]
        tree.insert(true_insertion, new_rotation, best_parent, best_heuristic, result["g"][c], result["insert_indices"][c])

        # ---- rewire the remaining neighbours through the new node (rrt_star.py:69-93)
        rest = [j for j in neighbor_indices if j != best_parent and tree.no_cycle(new_index, j)]
        if not rest:
            return
        re = None
        if parent_dependent:
            re = self._ftl_from_parents([new_index] * len(rest), [tree.nodes[j].insertion for j in rest],
                                        [tree.nodes[j].rotation for j in rest])
        for n, j in enumerate(rest):
            current = tree.nodes[j].heuristic
            if re is not None and re["ftl"][n] is None and hasattr(f, "create_from_magnitude"):
                candidate = f.create_from_magnitude(re["ftl_mag"][n], insertion_fraction=getattr(current, "insertion_fraction", None))
            else:
                candidate = f.create_from_old(current, follow_the_leader=re["ftl"][n] if re is not None else None)
            if candidate.test_cost_from_parent(best_heuristic) < current.get_cost():
                tree.swap_parents(j, new_index, candidate)
