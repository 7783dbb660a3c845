"""CPU: the collision oracle vs (a) the reference's own test literals at the FCL
boundary, (b) closed-form cases, (c) an independent NumPy convex search."""
import math

import numpy as np
import pytest

from conftest import FIXTURES
from oracle import oracle as orc
from oracle import np_collision as npc


def se3(p):
    g = np.eye(4)
    g[0:3, 3] = p
    return g


def test_reference_check_collision_literal():
    """test/test_collision.py:58-61: (-1, 0.01 +- 5e-7)."""
    env = orc.Env.from_json(FIXTURES / "init_box_mesh.json")
    obs, goal = env.check_collision([[se3([0, 0, 0]), se3([18.99, 0, 0])]], [1])
    assert obs == -1.0
    assert goal == pytest.approx(0.01, abs=5e-7)


def test_reference_cylinder_placement_literal():
    """test/test_collision.py:77-92: axis-aligned Cylinder obstacles collide with Sphere(1) at +-5."""
    env_prims = [orc._prim_from_json(o) for o in
                 __import__("json").loads((FIXTURES / "init_cylinder_test.json").read_text())["obstacles"]]
    for i in range(3):
        for j in (-1, 1):
            c = np.zeros(3)
            c[i] = 5 * j
            p = env_prims[i]
            axis = np.array(p.rot[:]).reshape(3, 3)[:, 2]
            d = orc.point_cylinder(c, np.array(p.center[:]), axis, 0.5 * p.dims[1], p.dims[0]) - 1.0
            assert d <= 0
            # and it does NOT collide along the other axes (cylinder radius 1 + sphere 1 < 5)
            k = (i + 1) % 3
            c2 = np.zeros(3)
            c2[k] = 5 * j
            assert orc.point_cylinder(c2, np.array(p.center[:]), axis, 0.5 * p.dims[1], p.dims[0]) - 1.0 > 0


def test_reference_build_tube_literals():
    """test/test_collision.py:24-56 (literal values, not the weak .all() asserts)."""
    assert orc.build_tube([[se3([1.0, 1.0, 1.0])]], [5]) == []
    t = orc.build_tube([[se3([0, 0, 30]), se3([0, 0, -5])]], [2])
    assert len(t) == 1 and t[0][1] == 35 and t[0][3].tolist() == [0, 1, 0, 0]
    t = orc.build_tube([[se3([0, 0, 0]), se3([0, 0, 3]), se3([0, 2, 3])]], [1])
    assert t[0][1] == 3 and t[0][2].tolist() == [0, 0, 1.5] and t[0][3].tolist() == [1, 0, 0, 0]
    assert t[1][1] == 2 and t[1][2].tolist() == [0, 1, 3]
    # the literal in the reference test is sqrt(2)/2-normalised (1,-1,0,0) scaled differently; check direction:
    q = t[1][3]
    np.testing.assert_allclose(q, np.array([1, -1, 0, 0]) / math.sqrt(2), atol=1e-15)


def test_closed_form_cases():
    c = np.zeros(3)
    a = np.array([0, 0, 1.0])
    h, r = 2.0, 1.0
    big = 50.0
    # parallel to the cap, 3 above it
    d = orc.cylinder_triangle(c, a, h, r, [-big, -big, 5], [big, -big, 5], [0, big, 5])
    assert d == pytest.approx(3.0, abs=1e-14)
    # facing the side, plane x = 4
    d = orc.cylinder_triangle(c, a, h, r, [4, -big, -big], [4, big, -big], [4, 0, big])
    assert d == pytest.approx(3.0, abs=1e-14)
    # tilted plane: nearest point is a rim point
    n = np.array([1.0, 0, 1.0]) / math.sqrt(2)
    p0 = n * 10
    u = np.array([0, 1.0, 0])
    v = np.cross(n, u)
    d = orc.cylinder_triangle(c, a, h, r, p0 - big * u - big * v, p0 + big * u - big * v, p0 + big * v)
    assert d == pytest.approx(10 - (2 + 1) / math.sqrt(2), abs=1e-13)
    # vertex nearest the rim
    d = orc.cylinder_triangle(c, a, h, r, [4, 0, 6], [40, 1, 60], [40, -1, 60])
    assert d == pytest.approx(5.0, abs=1e-13)
    # thin cylinder piercing a big triangle: all edges far away, still in collision
    d = orc.cylinder_triangle(c, a, h, 0.01, [-big, -big, 0.5], [big, -big, 0.5], [0, big, 0.5])
    assert d == 0.0
    # triangle completely inside the cylinder
    d = orc.cylinder_triangle(c, a, h, r, [0.1, 0, 0], [0, 0.1, 0], [0, 0, 0.1])
    assert d == 0.0
    # just touching the cap plane counts as contact; a hair above does not
    assert orc.cylinder_triangle(c, a, h, r, [-1, -1, 2.0], [1, -1, 2.0], [0, 1, 2.0]) == 0.0
    assert orc.cylinder_triangle(c, a, h, r, [-1, -1, 2.0 + 1e-9], [1, -1, 2.0 + 1e-9], [0, 1, 2.0 + 1e-9]) \
        == pytest.approx(1e-9, rel=1e-6)


def _random_pairs(rng, P):
    c = rng.normal(size=(P, 3)) * 3
    a = rng.normal(size=(P, 3))
    a /= np.linalg.norm(a, axis=1)[:, None]
    h = rng.uniform(0.2, 3.0, P)
    r = rng.uniform(0.2, 2.0, P)
    scale = rng.choice([0.3, 1.5, 8.0, 40.0], P)[:, None]
    off = rng.normal(size=(P, 3)) * rng.choice([0.5, 3.0, 10.0], P)[:, None]
    t = [c + off + rng.normal(size=(P, 3)) * scale for _ in range(3)]
    return c, a, h, r, t


def test_cylinder_triangle_vs_numpy_search():
    rng = np.random.default_rng(77)
    P = 1500
    c, a, h, r, t = _random_pairs(rng, P)
    want = npc.cylinder_triangle(c, a, h, r, *t)
    got = np.array([orc.cylinder_triangle(c[i], a[i], h[i], r[i], t[0][i], t[1][i], t[2][i]) for i in range(P)])
    pos = got > 0
    assert pos.sum() > 300 and (~pos).sum() > 100
    assert np.max(np.abs(got[pos] - want[pos])) < 1e-7
    assert np.all(want[~pos] < 1e-6)
    # the analytic result is never above the search (the search can only over-estimate a minimum)
    assert np.all(got <= want + 1e-12)


def test_segment_cylinder_vs_dense_sampling():
    rng = np.random.default_rng(5)
    for _ in range(300):
        c = rng.normal(size=3)
        a = rng.normal(size=3)
        a /= np.linalg.norm(a)
        h, r = rng.uniform(0.2, 2), rng.uniform(0.2, 2)
        p0, p1 = rng.normal(size=3) * 4, rng.normal(size=3) * 4
        s = np.linspace(0, 1, 20001)[:, None]
        x = p0 + s * (p1 - p0)
        dense = npc.point_cylinder(x, c, a, h, r).min()
        got = orc.segment_cylinder(p0, p1, c, a, h, r)
        assert got <= dense + 1e-12
        assert dense - got < 1e-6


def test_point_triangle_vs_sampling():
    rng = np.random.default_rng(6)
    for _ in range(200):
        t0, t1, t2 = rng.normal(size=(3, 3)) * 3
        x = rng.normal(size=3) * 4
        u = rng.random(40000)
        v = rng.random(40000)
        m = u + v > 1
        u[m], v[m] = 1 - u[m], 1 - v[m]
        pts = t0 + u[:, None] * (t1 - t0) + v[:, None] * (t2 - t0)
        dense = np.linalg.norm(pts - x, axis=1).min()
        got = orc.point_triangle(x, t0, t1, t2)
        assert got <= dense + 1e-12 and dense - got < 0.05


def test_tree_equals_brute_force_on_colon():
    bases, dof, q, lengths, dx = (["linear"] * 3, [2, 3, 3],
                                  [0.02, 0.001, 0.05, 0.0001, 0.002, 0.03, 0.0002, 0.0005], [33, 66, 99], 1)
    m = orc.make_model(bases, dof, q, lengths, dx)
    env = orc.Env.from_json(FIXTURES / "env_colon_straight_new.json")
    rng = np.random.default_rng(3)
    n_free = 0
    for _ in range(12):
        e = rng.integers(1, 9, 3)
        idx = [m.npts[n] - 1 - int(e[n]) for n in range(3)]
        th = rng.uniform(0, 2 * np.pi, 3)
        g = orc.solve_g(m, idx, th, full=False)
        a = env.check_collision(g, [1.2, 0.9, 0.6])
        b = env.check_collision(g, [1.2, 0.9, 0.6], brute=True)
        assert a == b
        n_free += a[0] > 0
    assert n_free > 0


def test_empty_obstacles_and_mesh_goal():
    env = orc.Env.from_json(FIXTURES / "env_ftl_goal.json")
    obs, goal = env.check_collision([[se3([0, 0, 0]), se3([5, 0, 0])]], [5])
    assert obs == orc.DBL_MAX  # collision_checker.py:89-95 with no obstacles: DistanceResult default
    assert goal > 0 or goal == -1.0


def test_cylinder_cylinder_gjk():
    """Obstacle Cylinder vs tube cylinder: closed-form cases and an independent SLSQP minimisation of the
    (convex) point-to-cylinder distance over the other cylinder."""
    from scipy.optimize import minimize
    z, y = np.array([0, 0, 1.0]), np.array([0, 1.0, 0])
    assert orc.cylinder_cylinder([0, 0, 0], z, 1, 1, [0, 0, 4], z, 2, 0.5) == pytest.approx(1.0, abs=1e-12)
    assert orc.cylinder_cylinder([0, 0, 0], z, 1, 1, [2.5, 0, 0.3], z, 1, 1) == pytest.approx(0.5, abs=1e-12)
    assert orc.cylinder_cylinder([0, 0, 0], z, 5, 1, [3, 0, 0], y, 5, 1) == pytest.approx(1.0, abs=1e-12)
    assert orc.cylinder_cylinder([0, 0, 0], z, 5, 1, [1.5, 0, 0], y, 5, 1) == 0.0
    assert orc.cylinder_cylinder([0, 0, 0], z, 1, 1, [3, 0, 3], z, 1, 1) == pytest.approx(np.sqrt(2), abs=1e-12)
    rng = np.random.default_rng(11)

    def pc(x, c, a, h, r):
        yv = x - c
        t = yv @ a
        return np.hypot(max(abs(t) - h, 0), max(np.linalg.norm(yv - t * a) - r, 0))

    n_pos = 0
    for _ in range(60):
        c1, c2 = rng.normal(size=3) * 2, rng.normal(size=3) * 3
        a1, a2 = rng.normal(size=3), rng.normal(size=3)
        a1, a2 = a1 / np.linalg.norm(a1), a2 / np.linalg.norm(a2)
        h1, r1, h2, r2 = rng.uniform(.3, 2), rng.uniform(.3, 1.5), rng.uniform(.3, 3), rng.uniform(.3, 2)
        d = orc.cylinder_cylinder(c1, a1, h1, r1, c2, a2, h2, r2)
        e1 = np.cross(a2, [1, 0, 0])
        e1 /= np.linalg.norm(e1)
        e2 = np.cross(a2, e1)
        f = lambda p: pc(c2 + p[0] * a2 + p[1] * e1 + p[2] * e2, c1, a1, h1, r1) ** 2  # noqa: E731
        best = np.inf
        for _s in range(5):
            p0 = np.array([rng.uniform(-h2, h2), rng.uniform(-r2, r2) / 2, rng.uniform(-r2, r2) / 2])
            res = minimize(f, p0, method="SLSQP", bounds=[(-h2, h2), (None, None), (None, None)],
                           constraints=[{"type": "ineq", "fun": lambda p: r2 * r2 - p[1] ** 2 - p[2] ** 2}],
                           options={"ftol": 1e-16, "maxiter": 500})
            best = min(best, np.sqrt(max(res.fun, 0)))
        if d > 0:
            n_pos += 1
            assert d <= best + 1e-7 and best - d < 1e-5  # SLSQP satisfies its constraint only to ~1e-8
        else:
            assert best < 1e-4
    assert n_pos > 25


def test_reference_cylinder_fixture_loads():
    """init_cylinder_test.json (test/test_collision.py:77-92) now goes through check_collision end to end."""
    env = orc.Env.from_json(FIXTURES / "init_cylinder_test.json")
    # tube (r = 0.5) along x at y = 3: 3 - 1 - 0.5 = 1.5 from the x- and y-aligned Cylinder(1, 10) obstacles
    obs, goal = env.check_collision([[se3([3.0, 3.0, 0.0]), se3([4.0, 3.0, 0.0])]], [0.5])
    assert obs == pytest.approx(1.5, abs=1e-12)
    # inside the z-aligned one
    obs, _ = env.check_collision([[se3([0.2, 0.3, 3.0]), se3([1.2, 0.3, 3.0])]], [0.5])
    assert obs == -1.0


@pytest.mark.parametrize("env_name", ["env_c3_colon_straight.json", "env_colon_straight_new.json", "env_colon_angled.json"])
def test_c_and_numpy_agree_on_the_colon_meshes(env_name):
    """The two independent implementations (the C case analysis and the NumPy nested convex search) on the REAL meshes:
    tube cylinders of random C3 configurations against the triangles of the posed colon files -- the triangles nearest to
    each cylinder (where the minimum is decided, contacts included) and random far ones.  With python-fcl unobtainable this
    agreement is what arbitrates the restated cylinder-triangle distance."""
    import json
    from conftest import FIXTURES
    spec = json.loads((FIXTURES / env_name).read_text())
    mesh = next(o for o in spec["obstacles"] if o.get("mesh"))
    tris = orc.read_stl(FIXTURES / mesh["filename"]).astype(np.float64) + np.asarray(mesh["transform"], np.float64)
    cent = tris.mean(1)
    from util import orc_model
    om = orc_model("c3_three_tube")
    rng = np.random.default_rng(2025)
    N = np.array([34, 67, 100])
    rad = [1.2, 0.9, 0.6]
    cs, as_, hs, rs, t0, t1, t2 = [], [], [], [], [], [], []
    for _ in range(40):
        e = rng.integers(2, 20, 3)
        g = orc.solve_g(om, (N - 1 - e).tolist(), rng.uniform(0, 2 * np.pi, 3).tolist(), full=False)
        for n, curve in enumerate(g):
            p = curve[:, 0:3, 3]
            k = int(rng.integers(0, len(p) - 1))
            f, t = p[k], p[k + 1]
            c, vec = (f + t) / 2, t - f
            h = np.linalg.norm(vec) / 2
            near = np.argsort(np.linalg.norm(cent - c, axis=1))[:6]
            far = rng.integers(0, len(tris), 2)
            for j in list(near) + list(far):
                cs.append(c); as_.append(vec / (2 * h)); hs.append(h); rs.append(rad[n])
                t0.append(tris[j, 0]); t1.append(tris[j, 1]); t2.append(tris[j, 2])
    cs, as_, hs, rs, t0, t1, t2 = (np.asarray(v) for v in (cs, as_, hs, rs, t0, t1, t2))
    want = npc.cylinder_triangle(cs, as_, hs, rs, t0, t1, t2)
    got = np.array([orc.cylinder_triangle(cs[i], as_[i], hs[i], rs[i], t0[i], t1[i], t2[i]) for i in range(len(cs))])
    pos = got > 0
    assert len(cs) == 960 and pos.sum() > 200
    assert np.max(np.abs(got[pos] - want[pos])) < 1e-7          # the search's own accuracy
    assert np.all(want[~pos] < 1e-6)                             # contacts are contacts for both
    assert np.all(got <= want + 1e-12)                           # a search can only over-estimate a minimum
