"""GPU: the planner mirror (ctrdapp_b200.solve) -- device nearest neighbours vs a NumPy scan, the reference's own
tree literals (test/test_solve.py), and whole RRT / RRT* runs: the sequential (batch_size 1) search driven through
the CUDA model + collision path must build the SAME tree as the same search driven through the CPU oracle."""
import numpy as np
import pytest

from conftest import FIXTURES
from util import gpu_model, orc_model

pytestmark = pytest.mark.gpu


def brute_knn(pts, q, w, topo, k):
    d = np.abs(pts[None, :, :] - q[:, None, :])
    circ = np.asarray(topo) == 2
    d[:, :, circ] = np.minimum(d[:, :, circ], np.abs(2 * np.pi - d[:, :, circ]))
    dist = np.sqrt(((d * w) ** 2).sum(-1))
    order = np.lexsort((np.broadcast_to(np.arange(len(pts)), dist.shape), dist), axis=1)[:, :k]
    return order, np.take_along_axis(dist, order, 1)


@pytest.mark.parametrize("T,N,k", [(1, 1, 1), (2, 3, 4), (3, 500, 16), (2, 4097, 8), (4, 70, 32)])
def test_knn_vs_numpy(T, N, k):
    import ctypes as C
    from ctrdapp_b200 import _lib as L
    rng = np.random.default_rng(T * 1000 + N)
    pts = np.empty((N, 2 * T))
    pts[:, 0::2] = rng.integers(0, 30, (N, T))           # many exact ties, like insertions on the delta_x grid
    pts[:, 1::2] = rng.uniform(0, 2 * np.pi, (N, T))
    if N > 10:
        pts[5] = pts[3]                                  # duplicate point: tie broken by index
    q = np.empty((64, 2 * T))
    q[:, 0::2] = rng.uniform(0, 30, (64, T))
    q[:, 1::2] = rng.uniform(0, 2 * np.pi, (64, T))
    q[0] = pts[min(3, N - 1)]
    w = np.ascontiguousarray([1.0, 1 / 7.5] * T)
    topo = np.ascontiguousarray([1, 2] * T, np.int32)
    idx = np.empty((64, k), np.int32)
    dist = np.empty((64, k))
    ctx = L.default_context()
    L.check(L.lib().ctrb_knn_batch(ctx.handle, 2 * T, N, L.ptr(pts), 64, L.ptr(q), L.ptr(w), L.ptr(topo), k, L.ptr(idx),
                                   L.ptr(dist), L.MEM_HOST))
    want_i, want_d = brute_knn(pts, q, w, topo, min(k, N))
    kk = min(k, N)
    # the device accumulates with fma: ties that NumPy sees as exact can differ in the last bit -> compare distances,
    # and indices wherever the distance gap to the next candidate is not a rounding artefact
    np.testing.assert_allclose(dist[:, :kk], want_d, rtol=1e-14, atol=1e-14)
    clear = np.ones_like(want_i, bool)
    clear[:, :-1] &= np.diff(want_d, axis=1) > 1e-12
    clear[:, 1:] &= np.diff(want_d, axis=1) > 1e-12
    assert np.array_equal(idx[:, :kk][clear], want_i[clear])
    assert np.all(idx[:, kk:] == -1) and np.all(np.isinf(dist[:, kk:]))
    assert idx[0, 0] == min(3, N - 1)                     # the duplicate at index 5 loses the tie to index 3


def test_reference_tree_literals():
    """test/test_solve.py:16-170 (scaling [1, 6]); the three find_all_nearest_neighbor literals that are consistent
    with the code (the first and the `[5,5,5,5] -> []` ones predate the "at least one neighbour" rule, :116-117)."""
    from ctrdapp_b200.heuristic import OnlyGoalDistance
    from ctrdapp_b200.solve import DynamicTree, Node
    h = [OnlyGoalDistance(c) for c in (10, 1, 2, 3)]
    simple_g, other_g = [[np.eye(4)]], [[np.eye(4)]]
    other_g[0][0][0:2, 3] = [0.5, 0.1]
    zeros, ones, twos, threes = [0.0, 0.0], [1.0, 1.0], [2.0, 2.0], [3.0, 3.0]
    tree = DynamicTree(2, zeros, zeros, h[0], simple_g, 10, [1, 6])
    assert tree.nearest_neighbor([0.0, 0.0, 0.0, 0.0]) == (zeros, zeros, 0, zeros)
    assert tree.find_all_nearest_neighbor([0.0, 0.0, 0.0, 0.0], 10) == [0]
    assert tree.find_all_nearest_neighbor([0.0, 0.0, 0.0, 0.0], 0.0001) == [0]
    tree.insert(ones, ones, 0, h[0], simple_g, [0, 0])
    tree.insert(twos, twos, 0, h[0], simple_g, [0, 0])
    tree.insert(threes, threes, 2, h[0], simple_g, [0, 0])
    assert len(tree.nodes) == 4 and tree.nodes[0].children == [1, 2] and tree.nodes[2].children == [3]

    final = DynamicTree(2, zeros, zeros, h[0], simple_g, 10, [1, 6])
    final.insert(ones, ones, 0, h[1], simple_g, [0, 0])
    final.insert(twos, twos, 1, h[2], simple_g, [0, 0])
    final.insert(threes, threes, 0, h[3], other_g, [1, 1])
    assert final.nearest_neighbor([0.0, 0.0, 0.0, 0.0]) == (zeros, zeros, 0, zeros)
    assert final.nearest_neighbor([1.0, 1.0, 1.0, 1.0]) == (ones, ones, 1, zeros)
    assert final.nearest_neighbor([0.8, 0.8, 0.7, 0.7]) == (ones, ones, 1, zeros)
    assert final.nearest_neighbor([2.5, 2.5, 2.49, 2.49]) == (twos, twos, 2, ones)
    assert final.nearest_neighbor([2.5, 2.5, 2.51, 2.51]) == (threes, threes, 3, zeros)
    assert final.find_all_nearest_neighbor([2.0, 1.0, 1.0, 1.0], 1.0001) == [1]
    assert sorted(final.find_all_nearest_neighbor([2.0, 2.0, 1.0, 1.0], 2.5)) == [0, 1, 2, 3]
    assert final.find_all_nearest_neighbor([2.0, 2.0, 1.0, 1.0], 2.5)[:2] == [1, 2]
    assert final.find_all_nearest_neighbor([2.0, 2.0, 0.0, 0.0], 2.5)[0] == 1
    assert sorted(final.find_all_nearest_neighbor([2.0, 2.0, 0.0, 0.0], 2.5)) == [0, 1, 2]
    assert final.get_costs(0) == [10] and final.get_costs(2) == [2, 1, 10] and final.get_costs(3) == [3, 10]
    assert final.get_tube_data(2) == ([twos, ones, zeros], [twos, ones, zeros], [[0, 0]] * 3)
    assert final.get_tube_data(3) == ([threes, zeros], [threes, zeros], [[1, 1], [0, 0]])
    assert final.get_tube_curves(3)[0] is other_g and len(final.get_tube_curves(2)) == 3
    with pytest.raises(ValueError):
        Node([0], [0], 2, h[0], simple_g, [0, 0])
    # test_no_cycle (:139-170)
    t = DynamicTree(1, [0.0], [0.0], h[0], simple_g, 10, [1, 6])
    for parent in (0, 1, 2, 0, 4, 4, 2, 7):
        t.insert([0.0], [0.0], parent, h[0], simple_g, [1])
    assert t.no_cycle(5, 3) and t.no_cycle(5, 2) and t.no_cycle(5, 1)
    assert not t.no_cycle(3, 2) and not t.no_cycle(8, 2) and not t.no_cycle(8, 1) and t.no_cycle(8, 4)
    with pytest.raises(ValueError):
        t.no_cycle(0, 3)


# ---------------------------------------------------------------- whole searches
CONF = {"tube_number": 3, "tube_radius": {"outer": [1.2, 0.9, 0.6], "inner": [1.0, 0.7, 0.4]}, "tube_lengths": [33, 66, 99],
        "iteration_number": 300, "step_bound": 7.5, "rotation_max": 0.5, "single_tube_control": False, "seed": 42,
        "goal_weight": 3.0, "only_tip": True, "insertion_weight": 0.25}
HDICT = {"square_obstacle_avg_plus_weighted_goal": ["goal_weight"], "only_goal_distance": [],
         "follow_the_leader": ["only_tip"], "follow_the_leader_w_insertion": ["only_tip", "insertion_weight"],
         "follow_the_leader_translation": ["only_tip"]}


class OracleModel:
    """The CPU oracle behind the reference's model interface (test infrastructure)."""

    def __init__(self, name):
        from oracle import oracle as orc
        from models_def import MODELS
        self.orc, self.m = orc, orc_model(name)
        _, dof, q, lengths, dx = MODELS[name]
        self.q, self.tube_num, self.max_tube_length, self.delta_x = q, len(lengths), lengths, dx

    def solve_integrate(self, delta_theta, delta_insertion, this_theta, this_insertion, prev_g, invert_insert=True,
                        need_g_out=True):
        return self.orc.solve_integrate(self.m, delta_theta, delta_insertion, this_theta, this_insertion, prev_g,
                                        invert_insert, need_g_out)


def _factory(name):
    from ctrdapp_b200.heuristic import create_heuristic_factory
    return create_heuristic_factory(dict(CONF, heuristic_type=name), HDICT)


def _same_tree(a, b):
    assert len(a.tree.nodes) == len(b.tree.nodes) > 10
    for x, y in zip(a.tree.nodes, b.tree.nodes):
        assert x.parent == y.parent and x.insert_indices == y.insert_indices
        assert x.insertion == y.insertion and x.rotation == y.rotation
        assert x.get_cost() == pytest.approx(y.get_cost(), rel=1e-9)
    assert a.tree.at_goal == b.tree.at_goal


@pytest.mark.parametrize("solver,heuristic,single", [("rrt", "square_obstacle_avg_plus_weighted_goal", False),
                                                     ("rrt", "follow_the_leader_w_insertion", True),
                                                     ("rrt_star", "follow_the_leader", False)])
def test_sequential_search_equals_oracle_driven_search(solver, heuristic, single):
    from oracle import oracle as orc
    from ctrdapp_b200.collision import CollisionChecker
    from ctrdapp_b200.solve import create_solver
    env = FIXTURES / "env_c3_colon_straight.json"
    conf = dict(CONF, solver_type=solver, single_tube_control=single, iteration_number=400)
    gpu = create_solver(gpu_model("c3_three_tube"), _factory(heuristic), CollisionChecker(env), conf)
    cpu = create_solver(OracleModel("c3_three_tube"), _factory(heuristic), orc.Env.from_json(env), conf)
    _same_tree(gpu, cpu)
    cost, best = gpu.get_best_cost()
    assert np.isfinite(cost) and 0 <= best < len(gpu.tree.nodes)


@pytest.mark.parametrize("solver", ["rrt", "rrt_star"])
def test_batched_search_invariants(solver, tmp_path):
    """batch_size 64: every stored node is collision-free according to the oracle, costs follow the parent chain,
    the saved tree reloads (TreeFromFile) with the same structure."""
    from oracle import oracle as orc
    from ctrdapp_b200.collision import CollisionChecker
    from ctrdapp_b200.solve import create_solver
    env = FIXTURES / "env_c3_colon_straight.json"
    conf = dict(CONF, solver_type=solver, batch_size=64, iteration_number=1024)
    m = gpu_model("c3_three_tube")
    s = create_solver(m, _factory("square_obstacle_avg_plus_weighted_goal"), CollisionChecker(env), conf)
    assert s.rounds == 16 and len(s.tree.nodes) > 50
    oenv = orc.Env.from_json(env)
    for i in range(1, len(s.tree.nodes), 7):
        node = s.tree.nodes[i]
        obs, goal = oenv.check_collision(node.g_curves, CONF["tube_radius"]["outer"])
        assert obs > 0
        assert abs(node.heuristic.store_obstacle_min - obs) < 1e-6
        chain = s.tree.get_index_list(i)
        own = [(1 / s.tree.nodes[j].heuristic.store_obstacle_min) ** 2 for j in chain[:-1]]
        h = node.heuristic
        assert h.generation == len(chain) - 1
        # running mean over the path; the root enters with generation 0, i.e. weight 0 (obstacles_and_goal.py:67-88)
        assert h.avg_obstacle_min == pytest.approx(np.mean(own), rel=1e-9)
    s.save_tree(tmp_path)
    lines = (tmp_path / "tree.txt").read_text().splitlines()
    assert len(lines) == len(s.tree.nodes) and lines[0].startswith("0 | 0, 0, 0 | 0, 0, 0 | None | ")
    # every line in the reference's format (dynamic_tree.py:328-355): "i | ins, .. | rot, .. | parent | cost | at_goal"
    import re
    num = r"-?\d+(?:\.\d+)?(?:e[+-]?\d+)?"
    line_re = re.compile(rf"^(\d+) \| {num}(?:, {num}){{2}} \| {num}(?:, {num}){{2}} \| (None|\d+) \| ({num}|inf|nan) \| (True|False)$")
    for k, ln in enumerate(lines):
        mt = line_re.match(ln)
        assert mt and int(mt.group(1)) == k, ln
        assert (mt.group(2) == "None") == (k == 0)       # (RRT* rewiring can hang a node under a later one)
        assert mt.group(2) == "None" or int(mt.group(2)) == s.tree.nodes[k].parent
        assert float(mt.group(3)) == pytest.approx(s.tree.nodes[k].heuristic.get_cost(), rel=1e-12, nan_ok=True)
        assert (mt.group(4) == "True") == (k in s.tree.at_goal)
    s.save_best_solution(tmp_path)
    sol = (tmp_path / "solution_path.txt").read_text().split("\n")   # rrt.py:184-198: "Q: ..", "Cost: ..", the index chain
    assert len(sol) == 3 and sol[0] == f"Q: {m.q}" and sol[1].startswith("Cost: ")
    cost, best = s.get_best_cost()
    assert float(sol[1][6:]) == pytest.approx(cost)
    chain = [int(v) for v in sol[2].split(", ")]
    assert chain[0] == 0 and chain[-1] == best and all(s.tree.nodes[b].parent == a for a, b in zip(chain, chain[1:]))
    # the reference's writer -> the reference's reader: integer tube lengths and delta_x give integer insertions
    # ("3, 6, 9", kinematic.py:74-78), which TreeFromFile parses back (tree_from_file.py:28) and re-solves
    import os
    from ctrdapp_b200.solve.tree_from_file import TreeFromFile
    assert all("." not in ln.split(" | ")[1] for ln in lines)
    run_dir = tmp_path / "output" / "reload"
    run_dir.mkdir(parents=True)
    s.save_tree(run_dir)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        t = TreeFromFile(m, _factory("square_obstacle_avg_plus_weighted_goal"), CollisionChecker(env),
                         dict(conf, run_identifier="reload", solver_type="tree_from_file"))
    finally:
        os.chdir(cwd)
    assert len(t.tree.nodes) == len(s.tree.nodes)
    key = lambda n: (tuple(float(v) for v in n.insertion), tuple(np.round(n.rotation, 12)))  # noqa: E731
    assert sorted(key(n) for n in t.tree.nodes) == sorted(key(n) for n in s.tree.nodes)
    by_key = {key(n): n for n in s.tree.nodes}
    for n in t.tree.nodes[1::9]:
        src = by_key[key(n)]
        assert [len(c) for c in n.g_curves] == [len(c) for c in src.g_curves]
        assert np.max(np.abs(np.concatenate(n.g_curves) - np.concatenate(src.g_curves))) < 1e-9
        if n.parent is not None:   # same parent configuration
            assert key(t.tree.nodes[n.parent]) == key(s.tree.nodes[src.parent])


def test_c4_follow_the_leader_goal_mesh():
    """BASELINE configs[3]: one tube r = 5 (test/configuration/config_ftl.yaml: linear dof 2, delta_x 2.5, step_bound
    7.5, RRT*, follow_the_leader_translation, only_tip) against the goal MESH follow_the_leader_goal_multiple.STL."""
    from ctrdapp_b200.collision import CollisionChecker
    from ctrdapp_b200.model.kinematic import Kinematic
    from ctrdapp_b200.solve import create_solver
    conf = {"tube_number": 1, "tube_radius": {"outer": [5.0], "inner": [4.0]}, "tube_lengths": [50], "delta_x": 2.5,
            "iteration_number": 800, "step_bound": 7.5, "rotation_max": 0.1745, "single_tube_control": True,
            "solver_type": "rrt_star", "heuristic_type": "follow_the_leader_translation", "only_tip": True,
            "batch_size": 32, "seed": 3}
    m = Kinematic(1, [np.array([0.03, 0.0005])], [2], [50], 2.5, ["linear"])
    from ctrdapp_b200.heuristic import create_heuristic_factory
    s = create_solver(m, create_heuristic_factory(conf, HDICT), CollisionChecker(FIXTURES / "env_ftl_goal.json"), conf)
    assert len(s.tree.nodes) == 801                       # no obstacles: every candidate is kept
    cost, best = s.get_best_cost()
    assert np.isfinite(cost) and cost >= 0
    # FTL costs are running sums of non-negative magnitudes along the path
    for i in range(1, len(s.tree.nodes), 13):
        node = s.tree.nodes[i]
        if any(v != 0.0 for v in node.insertion):
            assert node.get_cost() >= s.tree.nodes[node.parent].get_cost() - 1e-12 or node.heuristic.generation == 0


def test_nelder_mead_design_optimisation(tmp_path):
    """BASELINE configs[4] in miniature: Nelder-Mead over the 5 design parameters of a 2-tube robot, the objective of
    a design being the best cost of a batched RRT run against it (neldermead.py:101-135)."""
    from ctrdapp_b200.collision import CollisionChecker
    from ctrdapp_b200.heuristic import create_heuristic_factory
    from ctrdapp_b200.optimize import create_optimizer
    conf = {"optimizer_type": "nelder_mead", "optimizer_precision": 0.1, "optimize_iterations": 14,
            "solver_type": "rrt", "model_type": "kinematic", "heuristic_type": "square_obstacle_avg_plus_weighted_goal",
            "goal_weight": 3.0, "tube_number": 2, "tube_radius": {"outer": [0.9, 0.6], "inner": [0.7, 0.4]},
            "tube_lengths": [60, 120], "delta_x": 1, "strain_bases": ["linear", "linear"], "q_dof": [2, 3],
            "iteration_number": 256, "batch_size": 64, "step_bound": 7.5, "rotation_max": 0.5,
            "single_tube_control": False, "seed": 9}
    cc = CollisionChecker(FIXTURES / "env_boxes_multi.json")
    opt = create_optimizer(create_heuristic_factory(conf, HDICT), cc, [0.02, 0.001, 0.05, 0.0001, 0.002], conf)
    res = opt.find_min()
    assert res.num_function_eval >= 6 and res.best_solver is not None
    # one process: no speculation, the run is SciPy's algorithm evaluation for evaluation (the reference's)
    assert len(res.optimize_process) == res.num_function_eval and res.num_speculative_eval == 0
    assert res.owns_best_solver and res.best_rank == 0
    costs = [e["cost"] for e in res.optimize_process]
    assert min(costs) == pytest.approx(res.best_solver.get_best_cost()[0])
    assert len(res.best_q) == 5
    res.save_result(tmp_path)
    for name in ("tree.txt", "solution_path.txt", "optimize_result.txt", "optimize_process.txt"):
        assert (tmp_path / name).stat().st_size > 0
    assert (tmp_path / "optimize_result.txt").read_text().startswith("Optimizer Success: ")
    # the objective is deterministic (seeded planner): re-evaluating the best design gives the same cost
    assert opt.solver_heuristic(res.optimize_process[int(np.argmin(costs))]["q"]) == pytest.approx(min(costs))


def test_batched_rrt_star_tree_integrity():
    """batch_size 128 RRT* with a parent-dependent cost (one k-NN launch, one choose-parent batch and one rewiring batch
    per round): the tree stays a tree -- parent / children lists agree, every node reaches the root, FTL path costs are
    the running sums the heuristic defines -- and rewiring did happen."""
    from ctrdapp_b200.collision import CollisionChecker
    from ctrdapp_b200.model.kinematic import Kinematic
    from ctrdapp_b200.solve import create_solver
    from ctrdapp_b200.heuristic import create_heuristic_factory
    conf = {"tube_number": 1, "tube_radius": {"outer": [5.0], "inner": [4.0]}, "tube_lengths": [50], "delta_x": 2.5,
            "iteration_number": 1536, "step_bound": 7.5, "rotation_max": 0.1745, "single_tube_control": False,
            "solver_type": "rrt_star", "heuristic_type": "follow_the_leader_translation", "only_tip": True,
            "batch_size": 128, "seed": 9}
    m = Kinematic(1, [np.array([0.02, 0.0005])], [2], [50], 2.5, ["linear"])
    s = create_solver(m, create_heuristic_factory(conf, HDICT), CollisionChecker(FIXTURES / "env_ftl_goal.json"), conf)
    nodes = s.tree.nodes
    assert len(nodes) == 1537 and s.rounds == 12
    rewired = 0
    for i in range(1, len(nodes)):
        p = nodes[i].parent
        assert 0 <= p < len(nodes) and i in nodes[p].children
        chain = s.tree.get_index_list(i)
        assert chain[-1] == 0 and len(set(chain)) == len(chain)
        rewired += p > i                      # a parent younger than its child can only come from a rewire
    assert rewired > 0
    assert sum(len(n.children) for n in nodes) == len(nodes) - 1
