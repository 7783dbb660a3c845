"""CPU: host logic of the planner mirror (ctrdapp_b200.solve.step, ctrdapp_b200.heuristic) against golden vectors
produced by the reference itself (tests/golden/make_golden_solve.py), plus the tree bookkeeping literals of the
reference's test/test_solve.py that do not need a device."""
import io
import json
from contextlib import redirect_stdout

import numpy as np
import pytest

from conftest import GOLDEN
from ctrdapp_b200.heuristic.heuristic_factory import create_heuristic_factory
from ctrdapp_b200.solve.step import (delta_rotation_batch, get_delta_rotation, get_single_tube_value, step, step_batch,
                                     step_rotation, step_rotation_batch)

HDICT = {"square_obstacle_avg_plus_weighted_goal": ["goal_weight"], "only_goal_distance": [],
         "follow_the_leader": ["only_tip"], "follow_the_leader_w_insertion": ["only_tip", "insertion_weight"],
         "follow_the_leader_translation": ["only_tip"]}


@pytest.fixture(scope="module")
def gold():
    return json.loads((GOLDEN / "solve.json").read_text())


def test_delta_rotation(gold):
    a = np.array(gold["delta_rotation"])
    for prev, to, want in a:
        assert get_delta_rotation(prev, to) == want
    np.testing.assert_array_equal(delta_rotation_batch(a[:, 0], a[:, 1]), a[:, 2])


def test_step_and_step_rotation(gold):
    for c in gold["step"]:
        np.testing.assert_allclose(step(c["prev"], c["ctrl"], c["bound"]), c["out"], rtol=1e-15, atol=0)
    for c in gold["step_rotation"]:
        fin, dl = step_rotation(c["prev"], c["ctrl"], c["rmax"])
        np.testing.assert_allclose(fin, c["final"], rtol=1e-15, atol=1e-15)
        np.testing.assert_allclose(dl, c["delta"], rtol=1e-15, atol=1e-15)
    # stacked candidates == one at a time
    same = [c for c in gold["step"] if len(c["prev"]) == 2 and c["bound"] == 10.0]
    out = step_batch(np.array([c["prev"] for c in same]), np.array([c["ctrl"][::2] for c in same]), 10.0)
    np.testing.assert_allclose(out, [c["out"] for c in same], rtol=1e-15)
    same = [c for c in gold["step_rotation"] if len(c["prev"]) == 3 and c["rmax"] == 0.3]
    fin, dl = step_rotation_batch(np.array([c["prev"] for c in same]), np.array([c["ctrl"][1::2] for c in same]), 0.3)
    np.testing.assert_allclose(fin, [c["final"] for c in same], rtol=1e-15, atol=1e-15)
    np.testing.assert_allclose(dl, [c["delta"] for c in same], rtol=1e-15, atol=1e-15)
    with pytest.raises(ValueError):
        step([1.0, 2.0], [1.0, 0.0], 1.0)


def test_single_tube_value(gold):
    for c in gold["single_tube"]:
        out, tube = get_single_tube_value(list(c["ctrl"]), c["ins"], c["rot"], c["parent"], 0.8, c["r"])
        assert tube == c["tube"] and list(out) == c["out"]
    with pytest.raises(ValueError):
        get_single_tube_value([0] * 4, [1.0, 1.0], [0, 0], [0.0, 0.0], 0.8, 0.1)


def test_heuristics_follow_the_reference(gold):
    for case in gold["heuristic"]:
        f = create_heuristic_factory(case["conf"], HDICT)
        nodes = []
        for d, e in enumerate(case["chain"]):
            kw = dict(e["kw"])
            kw["follow_the_leader"] = [[np.array(a) for a in t] for t in kw["follow_the_leader"]]
            h = f.create(**kw)
            assert h.get_cost() == pytest.approx(e["own_cost"], rel=1e-15)
            if nodes:
                with redirect_stdout(io.StringIO()):
                    assert h.test_cost_from_parent(nodes[-1]) == pytest.approx(e["test_cost"], rel=1e-15)
                    h.calculate_cost_from_parent(nodes[-1], init_insertion=(d == 3 and case["only_tip"]))
                assert h.get_cost() == pytest.approx(e["cost_after"], rel=1e-15)
                assert getattr(h, "generation", None) == e["generation"]
            nodes.append(h)
        with redirect_stdout(io.StringIO()):
            old = f.create_from_old(nodes[3], follow_the_leader=[[np.array(a) for a in t] for t in case["rewire_ftl"]])
            assert old.test_cost_from_parent(nodes[1]) == pytest.approx(case["rewire_test_cost"], rel=1e-15)
            old.calculate_cost_from_parent(nodes[1], reset=True)
        assert old.get_cost() == pytest.approx(case["rewire_cost_after"], rel=1e-15)
    with pytest.raises(UserWarning):
        create_heuristic_factory({"heuristic_type": "nope"}, {"nope": []})


def test_reference_heuristic_literals():
    """test/test_heuristic.py:31-71 (the passing SOAPWG arithmetic)."""
    from ctrdapp_b200.heuristic import OnlyGoalDistance, SquareObstacleAvgPlusWeightedGoal
    a = SquareObstacleAvgPlusWeightedGoal(2, 0.5, 10)
    assert a.get_cost() == 2 * 10 + 4.0
    b = SquareObstacleAvgPlusWeightedGoal(2, 0.25, 4)
    b.calculate_cost_from_parent(a)
    assert b.generation == 1 and b.avg_obstacle_min == pytest.approx((0 * 4.0 + 16.0) / 1)
    c = SquareObstacleAvgPlusWeightedGoal(2, 1.0, 1)
    assert c.test_cost_from_parent(b) == pytest.approx(2 * 1 + (1 * 16.0 + 1.0) / 2)
    assert c.avg_obstacle_min == 1.0  # test_cost does not mutate
    g = OnlyGoalDistance(7.5)
    g.calculate_cost_from_parent(g)
    assert g.get_cost() == 7.5


def test_stl_loader_binary_and_ascii():
    """init_collision.py:59-88 through the mirror's parse_mesh: SURVEY App. C triangle counts and bounding boxes;
    sphere_100mm.STL is ASCII (CRLF), the colon files are binary although their header starts with 'solid'."""
    from conftest import FIXTURES
    from ctrdapp_b200.collision.init_collision import parse_mesh
    from oracle import oracle as orc
    expect = {"colon_straight.STL": (11492, [-312.90, -154.22, 214.47], [-234.00, 0.88, 264.35]),
              "sphere_100mm.STL": (2400, [-60.5, -49.6, -39.4], [17.9, 29.0, 39.4]),
              "mm_simple.STL": (12, [0, 0, 0], [1, 2, 3]),
              "follow_the_leader_goal_multiple.STL": (5268, [-7, -34, -34], [46, 34, 34])}
    for name, (ntri, lo, hi) in expect.items():
        v, tris = parse_mesh(FIXTURES / name)
        assert v.dtype == np.float32 and len(tris) == ntri and v.shape == (3 * ntri, 3)
        np.testing.assert_allclose(v.min(0), lo, atol=0.06)
        np.testing.assert_allclose(v.max(0), hi, atol=0.06)
        np.testing.assert_array_equal(orc.read_stl(FIXTURES / name).reshape(-1, 3), v.astype(np.float64))


def test_static_degree_matches_the_reference_probe():
    """static `degree` (static_bases.py:61-74): degree 0 / 1 fail in the reference with an IndexError (its hard-coded
    column offsets 3 and 6), degree > 2 runs on a malformed basis (profiles/r2_probe_static_degree.txt): the mirror raises
    the same exception type for the former and NotImplementedError for the latter, before touching the device."""
    import pathlib
    from ctrdapp_b200.model.static import Static, check_degree
    check_degree(2)
    for d in (0, 1):
        with pytest.raises(IndexError, match="out of bounds for axis 1"):
            check_degree(d)
        with pytest.raises(IndexError):
            Static(2, [np.zeros(2), np.zeros(3)], [2, 3], [120, 120], 1, ["linear", "linear"], {"outer": [0.9, 0.6], "inner": [0.7, 0.4]}, d)
    with pytest.raises(NotImplementedError, match="malformed"):
        check_degree(3)
    probe = (pathlib.Path(__file__).resolve().parent.parent / "profiles" / "r2_probe_static_degree.txt").read_text()
    assert probe.count("FAILS: IndexError") == 2 and "all-zero columns: [10, 11]" in probe
