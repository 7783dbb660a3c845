"""GPU parity: tube-vs-environment distance / collision / cost kernels vs the CPU oracle and
the reference's own literals at the FCL boundary."""
import numpy as np
import pytest

from conftest import FIXTURES
from util import gpu_model, orc_model, random_configs

pytestmark = pytest.mark.gpu

RAD3 = [1.2, 0.9, 0.6]
BAND = 1e-6  # mm; north-star boundary band for the collision bit and distance tolerance


def se3(p):
    g = np.eye(4)
    g[0:3, 3] = p
    return g


def _compare(obs, goal, o_obs, o_goal):
    """Parity contract of SURVEY 8c: bit exact outside the band, distances within 1e-6 mm."""
    in_band = (np.abs(o_obs) < BAND) & (o_obs > 0)
    bit, o_bit = obs < 0, o_obs < 0
    assert np.array_equal(bit[~in_band], o_bit[~in_band])
    free = ~bit & ~o_bit
    assert np.all(obs[bit] == -1.0)
    if free.any():
        assert np.max(np.abs(obs[free] - o_obs[free])) < BAND
    gb, o_gb = goal < 0, o_goal < 0
    assert np.array_equal(gb, o_gb)
    assert np.all(goal[gb] == -1.0)
    if (~gb).any():
        assert np.max(np.abs(goal[~gb] - o_goal[~gb])) < BAND
    return int(in_band.sum())


def test_reference_literals():
    """test/test_collision.py:58-61 and :24-56 through the mirrored CollisionChecker."""
    from ctrdapp_b200.collision import CollisionChecker
    cc = CollisionChecker(FIXTURES / "init_box_mesh.json")
    obs, goal = cc.check_collision([[se3([0, 0, 0]), se3([18.99, 0, 0])]], [1])
    assert obs == -1
    assert goal == pytest.approx(0.01, abs=5e-7)
    # a single node builds no cylinder (test_build_tube: []) -> nothing can collide
    obs, goal = cc.check_collision([[se3([10.0, 10.0, 10.0])]], [5])
    assert obs == np.finfo(np.float64).max


@pytest.mark.parametrize("env_name,lo,hi", [("env_colon_straight_new.json", 1, 8), ("env_colon_straight_new.json", 1, 33),
                                            ("env_c3_colon_straight.json", 1, 8), ("env_c3_colon_straight.json", 1, 20),
                                            ("env_colon_angled.json", 1, 12)])
def test_fused_eval_vs_oracle(env_name, lo, hi):
    from oracle import oracle as orc
    from ctrdapp_b200.collision import CollisionChecker
    m, om = gpu_model("c3_three_tube"), orc_model("c3_three_tube")
    cc = CollisionChecker(FIXTURES / env_name)
    oenv = orc.Env.from_json(FIXTURES / env_name)
    rng = np.random.default_rng(1003)
    B = 3000
    idx, theta = random_configs(m.num_discrete_points, B, rng, lo, hi)
    obs, goal, cost, coll = cc.evaluate_batch(m, idx, theta, RAD3, goal_weight=3.0)
    o_obs, o_goal, o_cost = oenv.eval_batch(om, idx, theta, RAD3, 3.0)
    _compare(obs, goal, o_obs, o_goal)
    assert np.array_equal(coll.astype(bool), obs < 0)
    free = obs > 0
    assert 0 < free.sum()
    assert np.all(np.isnan(cost[~free]))
    np.testing.assert_allclose(cost[free], o_cost[free], rtol=1e-9)  # (1/d)^2 amplification near contact
    # deterministic and order independent
    perm = rng.permutation(B)
    obs2, goal2, cost2, _ = cc.evaluate_batch(m, idx[perm], theta[perm], RAD3, goal_weight=3.0)
    assert np.array_equal(obs2, obs[perm]) and np.array_equal(goal2, goal[perm])


def test_check_collision_batch_given_curves():
    from oracle import oracle as orc
    from ctrdapp_b200.collision import CollisionChecker
    m, om = gpu_model("config_yaml"), orc_model("config_yaml")
    cc = CollisionChecker(FIXTURES / "env_boxes_multi.json")
    oenv = orc.Env.from_json(FIXTURES / "env_boxes_multi.json")
    rng = np.random.default_rng(1001)
    B = 256
    idx, theta = random_configs(m.num_discrete_points, B, rng)
    g, off = m.solve_g_batch(idx, theta)
    obs, goal = cc.check_collision_batch(g, off, [0.9, 0.6])
    o_obs = np.zeros(B)
    o_goal = np.zeros(B)
    for b in range(B):
        curve = [g[off[b * 2 + n]:off[b * 2 + n + 1]] for n in range(2)]
        o_obs[b], o_goal[b] = oenv.check_collision(curve, [0.9, 0.6])
    _compare(obs, goal, o_obs, o_goal)
    assert (obs > 0).sum() > 10 and (obs < 0).sum() > 0


def test_c3_survey_pose_all_collide():
    """SURVEY 8d's pose of colon_straight.STL puts a wall within 3.4 mm of the robot base:
    every configuration with >= 1 mm exposed per tube collides (first-contact exit regime)."""
    from ctrdapp_b200.collision import CollisionChecker
    m = gpu_model("c3_three_tube")
    cc = CollisionChecker(FIXTURES / "env_c3_survey_pose.json")
    rng = np.random.default_rng(1003)
    idx, theta = random_configs(m.num_discrete_points, 20000, rng, 1, 33)
    obs, goal, cost, coll = cc.evaluate_batch(m, idx, theta, RAD3, goal_weight=3.0)
    assert (obs == -1.0).mean() > 0.999 and np.all(goal > 0)


def test_c1_box_at_base_all_collide():
    """BASELINE configs[0] environment: in_simple.STL has a corner at the robot base."""
    from ctrdapp_b200.collision import CollisionChecker
    m = gpu_model("config_yaml")
    cc = CollisionChecker(FIXTURES / "env_c1_in_simple.json")
    rng = np.random.default_rng(1001)
    idx = rng.integers(0, 120, (1024, 2)).astype(np.int32)
    theta = rng.uniform(0, 2 * np.pi, (1024, 2))
    obs, goal, cost, coll = cc.evaluate_batch(m, idx, theta, [0.9, 0.6], goal_weight=3.0)
    assert np.all(obs == -1.0) and np.all(coll == 1) and np.all(np.isnan(cost))


def test_primitive_obstacles_and_goals():
    from oracle import oracle as orc
    from ctrdapp_b200 import _lib as L
    from ctrdapp_b200.collision import CollisionChecker
    from ctrdapp_b200.collision.init_collision import EnvObject

    def prim(shape, center, dims, rot=np.eye(3)):
        p, q = L.Prim(), orc.OrcPrim()
        for o in (p, q):
            o.shape = shape
            o.center[:] = list(map(float, center))
            o.dims[:] = list(map(float, dims))
            o.rot[:] = np.asarray(rot, float).ravel().tolist()
        return p, q

    m, om = gpu_model("c3_three_tube"), orc_model("c3_three_tube")
    rng = np.random.default_rng(4)
    idx, theta = random_configs(m.num_discrete_points, 500, rng, 1, 30)
    rot = orc.rotate_align(np.array([1.0, 2.0, 2.0]) / 3.0)
    obstacles = [prim(L.SHAPE_SPHERE, (30, 8, -3), (4, 0, 0)), prim(L.SHAPE_BOX, (25, -12, 6), (6, 8, 10)),
                 prim(L.SHAPE_SPHERE, (45, 0, 12), (6, 0, 0)),
                 prim(L.SHAPE_CYLINDER, (20, 9, 9), (2.5, 14, 0), orc.rotate_align(np.array([2.0, -1.0, 2.0]) / 3.0))]
    goals = [prim(L.SHAPE_SPHERE, (60, 5, 5), (5, 0, 0)), prim(L.SHAPE_BOX, (50, 0, 0), (10, 12, 14)),
             prim(L.SHAPE_CYLINDER, (40, 10, 5), (4, 20, 0), rot)]
    for gp, go in goals:
        cc = CollisionChecker.from_objects([EnvObject(p.shape, prim=p) for p, _ in obstacles],
                                           EnvObject(gp.shape, prim=gp))
        oenv = orc.Env(None, [q for _, q in obstacles], go)
        obs, goal, _, _ = cc.evaluate_batch(m, idx, theta, RAD3)
        o_obs, o_goal, _ = oenv.eval_batch(om, idx, theta, RAD3, 1.0)
        _compare(obs, goal, o_obs, o_goal)
        assert (obs > 0).sum() > 20 and (obs < 0).sum() > 20


def test_mesh_goal_no_obstacles():
    """C4 environment: goal mesh follow_the_leader_goal_multiple.STL, empty obstacle list."""
    from oracle import oracle as orc
    from ctrdapp_b200.collision import CollisionChecker
    from ctrdapp_b200.model.kinematic import Kinematic
    m = Kinematic(1, [np.array([0.02, 0.001])], [2], [50], 2.5, ["linear"])
    om = orc.make_model(["linear"], [2], [0.02, 0.001], [50], 2.5)
    cc = CollisionChecker(FIXTURES / "env_ftl_goal.json")
    oenv = orc.Env.from_json(FIXTURES / "env_ftl_goal.json")
    rng = np.random.default_rng(8)
    idx, theta = random_configs(m.num_discrete_points, 400, rng)
    obs, goal, cost, _ = cc.evaluate_batch(m, idx, theta, [5.0], heuristic_type="only_goal_distance")
    o_obs, o_goal, _ = oenv.eval_batch(om, idx, theta, [5.0], 1.0)
    assert np.all(obs == np.finfo(np.float64).max)
    _compare(obs, goal, o_obs, o_goal)
    np.testing.assert_allclose(cost, np.maximum(goal, 0.0))


def test_unsupported_and_edge_cases():
    from ctrdapp_b200.collision import CollisionChecker
    # the reference's own cylinder fixture (test/test_collision.py:77-92): axis-aligned Cylinder(1, 10) obstacles
    from oracle import oracle as orc
    cyl = CollisionChecker(FIXTURES / "init_cylinder_test.json")
    oenv = orc.Env.from_json(FIXTURES / "init_cylinder_test.json")
    rng = np.random.default_rng(2)
    for _ in range(40):
        p0 = rng.uniform(-7, 7, 3)
        d = rng.normal(size=3)
        p1 = p0 + d / np.linalg.norm(d)
        curve = [[se3(p0), se3(p1), se3(p1 + (p1 - p0))]]
        got, want = cyl.check_collision(curve, [0.5]), oenv.check_collision(curve, [0.5])
        assert (got[0] < 0) == (want[0] < 0) and abs(got[0] - want[0]) < 1e-6 and abs(got[1] - want[1]) < 1e-6
    assert cyl.check_collision([[se3([3.0, 3.0, 0.0]), se3([4.0, 3.0, 0.0])]], [0.5])[0] == pytest.approx(1.5, abs=1e-12)
    m = gpu_model("c3_three_tube")
    cc = CollisionChecker(FIXTURES / "env_colon_straight_new.json")
    out = cc.evaluate_batch(m, np.zeros((0, 3), np.int32), np.zeros((0, 3)), RAD3)
    assert all(len(o) == 0 for o in out)
    with pytest.raises(ValueError):
        cc.evaluate_batch(m, [[0, 0, 500]], [[0.0, 0.0, 0.0]], RAD3)
    st = cc.stats()
    assert st["triangles"] == 11455 and st["bvh_depth"] < 44


def test_large_batch_properties():
    """Size-independent checks at 10^6 configurations (BASELINE configs[2] at a tenth of its size):
    the collision bit agrees with the sign, the cost follows from the two distances, a random
    sub-sample matches the oracle."""
    from oracle import oracle as orc
    from ctrdapp_b200.collision import CollisionChecker
    m, om = gpu_model("c3_three_tube"), orc_model("c3_three_tube")
    cc = CollisionChecker(FIXTURES / "env_c3_colon_straight.json")
    rng = np.random.default_rng(1003)
    B = 1_000_000
    idx, theta = random_configs(m.num_discrete_points, B, rng, 1, 8)
    obs, goal, cost, coll = cc.evaluate_batch(m, idx, theta, RAD3, goal_weight=3.0)
    assert np.array_equal(coll == 1, obs == -1.0)
    assert np.all((obs == -1.0) | (obs > 0))
    free = obs > 0
    assert 0.1 < free.mean() < 0.6
    np.testing.assert_allclose(cost[free], 3.0 * np.maximum(goal[free], 0) + 1.0 / obs[free] ** 2, rtol=1e-12)
    sub = rng.choice(B, 1500, replace=False)
    oenv = orc.Env.from_json(FIXTURES / "env_c3_colon_straight.json")
    o_obs, o_goal, _ = oenv.eval_batch(om, idx[sub], theta[sub], RAD3, 3.0)
    _compare(obs[sub], goal[sub], o_obs, o_goal)


@pytest.mark.parametrize("B", [1, 31, 33, 1000, 2369, 100_003])
def test_grouped_and_one_warp_per_config_kernels_agree(B):
    """chain_eval_group_kernel (several configurations per warp; chunk size chosen from the batch size) and
    chain_eval_kernel (one warp per configuration; ctx option eval_group = 0) are two schedules of the same exact query:
    identical bits, distances and costs on ragged batch sizes, mixed short and long configurations included."""
    from ctrdapp_b200.collision import CollisionChecker
    m = gpu_model("c3_three_tube")
    cc = CollisionChecker(FIXTURES / "env_c3_colon_straight.json")
    rng = np.random.default_rng(B)
    idx, theta = random_configs(m.num_discrete_points, B, rng, 1, 8)
    long_ones = rng.random(B) < 0.2            # a fifth of the batch fully random (up to 198 cylinders)
    idx2, _ = random_configs(m.num_discrete_points, B, rng)
    idx[long_ones] = idx2[long_ones]
    a = cc.evaluate_batch(m, idx, theta, RAD3, goal_weight=3.0)
    with cc.ctx.options(eval_group=0):
        b = cc.evaluate_batch(m, idx, theta, RAD3, goal_weight=3.0)
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x, y)


def test_model_larger_than_the_node_pool_takes_the_one_warp_kernel():
    """A configuration with more nodes than the grouped kernel's per-warp pool (256) falls back to one warp per
    configuration; results against the oracle as for any other model."""
    from oracle import oracle as orc
    from ctrdapp_b200.collision import CollisionChecker
    from ctrdapp_b200.model.kinematic import Kinematic
    from util import split_q
    bases, dof = ["linear"] * 3, [2, 3, 3]
    q = [0.01, 0.0005, 0.02, 0.0001, 0.001, 0.015, 0.0002, 0.0003]
    lengths = [60, 120, 180]                   # 61 + 121 + 181 = 363 nodes
    m = Kinematic(3, split_q(q, dof), dof, lengths, 1, bases)
    om = orc.make_model(bases, dof, q, lengths, 1)
    cc = CollisionChecker(FIXTURES / "env_boxes_multi.json")
    oenv = orc.Env.from_json(FIXTURES / "env_boxes_multi.json")
    rng = np.random.default_rng(11)
    idx, theta = random_configs(m.num_discrete_points, 200, rng)
    obs, goal, cost, coll = cc.evaluate_batch(m, idx, theta, RAD3, goal_weight=3.0)
    o_obs, o_goal, _ = oenv.eval_batch(om, idx, theta, RAD3, 3.0)
    _compare(obs, goal, o_obs, o_goal)


def test_host_memory_pipeline_equals_single_shot():
    """Host-memory calls on >= 2^19 configurations run as a chunked copy / compute pipeline (two device buffer sets,
    two streams): the results are those of the same configurations evaluated in pieces small enough to go in one
    shot; a bad index in a late chunk is still a ValueError and leaves the context usable."""
    from ctrdapp_b200.collision import CollisionChecker
    m = gpu_model("c3_three_tube")
    cc = CollisionChecker(FIXTURES / "env_c3_colon_straight.json")
    rng = np.random.default_rng(77)
    B = 600_017                              # 5 chunks of 131 072 (the last one ragged)
    idx, theta = random_configs(m.num_discrete_points, B, rng, 1, 8)
    whole = cc.evaluate_batch(m, idx, theta, RAD3, goal_weight=3.0)
    cut = [0, 200_000, 400_000, B]
    parts = [cc.evaluate_batch(m, idx[a:b], theta[a:b], RAD3, goal_weight=3.0) for a, b in zip(cut[:-1], cut[1:])]
    for k in range(4):
        np.testing.assert_array_equal(whole[k], np.concatenate([p[k] for p in parts]))
    bad = idx.copy()
    bad[B - 5, 2] = 100                      # npts[2] = 100: out of range, last chunk
    with pytest.raises(ValueError):
        cc.evaluate_batch(m, bad, theta, RAD3, goal_weight=3.0)
    again = cc.evaluate_batch(m, idx[:1000], theta[:1000], RAD3, goal_weight=3.0)
    np.testing.assert_array_equal(again[0], whole[0][:1000])


def test_long_cylinders_vs_oracle():
    """Curves with very long segments (the static model's one-step sections produce them: metres between two nodes).
    The grouped kernel sets such cylinders aside and culls them by their axis (clip to the mesh box, slab test,
    segment-triangle prefilter); the answer must stay the exact one: piercing, grazing, free and far-away cases."""
    from oracle import oracle as orc
    from ctrdapp_b200.collision import CollisionChecker
    rng = np.random.default_rng(5)
    for env_name in ("env_c3_colon_straight.json", "env_colon_angled.json", "env_boxes_multi.json"):
        cc = CollisionChecker(FIXTURES / env_name)
        oenv = orc.Env.from_json(FIXTURES / env_name)
        curves, flat, off = [], [], [0]
        for b in range(160):
            n_nodes = int(rng.integers(2, 7))
            p = rng.uniform(-1.0, 1.0, 3)
            nodes = [se3(p)]
            for k in range(n_nodes - 1):
                d = rng.normal(size=3)
                d /= np.linalg.norm(d)
                step = rng.choice([1.0, 2.5, 4.0, 9.0, 30.0, 150.0, 4000.0, 3.0e6, 1.6e7])  # up to 16 000 km (seen in C3 batches)
                if b % 4 == 0:
                    step = 1.0 if k else rng.choice([6.0, 12.0])      # short-ish, mostly free inside the lumen
                p = p + step * d
                nodes.append(se3(p))
            curves.append([nodes])
            flat.extend(nodes)
            off.append(off[-1] + n_nodes)
        g = np.ascontiguousarray(np.stack(flat))
        obs, goal = cc.check_collision_batch(g, np.asarray(off, np.int64), [0.8])
        o = np.array([oenv.check_collision(c, [0.8]) for c in curves])
        # distances to far-away geometry: relative 1e-9 instead of the absolute micrometre band
        bit, o_bit = obs < 0, o[:, 0] < 0
        band = (np.abs(o[:, 0]) < BAND) & (o[:, 0] > 0)
        assert np.array_equal(bit[~band], o_bit[~band]), env_name
        free = ~bit & ~o_bit
        assert free.sum() >= 10 and (bit.sum() >= 3 or "boxes" in env_name), (env_name, free.sum(), bit.sum())
        assert np.all(np.abs(obs[free] - o[free, 0]) <= np.maximum(BAND, 1e-9 * o[free, 0])), env_name
