"""Golden vectors for the planner's host logic, produced by IMPORTING THE UNMODIFIED REFERENCE in the dev
container (ctrdapp.solve.step and ctrdapp.heuristic.* import without pympnn / fcl; the tree and the RRT classes
do not).  Run:  PYTHONPATH=/root/reference python tests/golden/make_golden_solve.py   -> tests/golden/solve.json
"""
import io
import json
import pathlib
import random
from contextlib import redirect_stdout
from math import pi

import numpy as np

from ctrdapp.heuristic.heuristic_factory import create_heuristic_factory
from ctrdapp.solve.step import get_delta_rotation, get_single_tube_value, step, step_rotation

OUT = pathlib.Path(__file__).resolve().parent / "solve.json"
rng = random.Random(20261019)
cases = {"delta_rotation": [], "step": [], "step_rotation": [], "single_tube": [], "heuristic": []}

for _ in range(200):
    a, b = rng.random() * 2 * pi, rng.random() * 2 * pi
    if rng.random() < 0.1:
        b = a
    cases["delta_rotation"].append([a, b, get_delta_rotation(a, b)])

for _ in range(200):
    T = rng.choice([1, 2, 3])
    prev = [rng.random() * 40 for _ in range(T)]
    rot_prev = [rng.random() * 2 * pi for _ in range(T)]
    ctrl = []
    for t in range(T):
        ctrl += [rng.random() * 40 if rng.random() < 0.8 else prev[t] + rng.random(), rng.random() * 2 * pi]
    bound = rng.choice([0.5, 2.7386, 10.0])
    rmax = rng.choice([0.05, 0.3, 5.0])
    cases["step"].append({"prev": prev, "ctrl": ctrl, "bound": bound, "out": list(step(prev, ctrl, bound))})
    fin, dl = step_rotation(rot_prev, ctrl, rmax)
    cases["step_rotation"].append({"prev": rot_prev, "ctrl": ctrl, "rmax": rmax, "final": list(fin), "delta": list(dl)})

for _ in range(100):
    T = rng.choice([1, 2, 3, 4])
    ins_n = [float(rng.randint(0, 20)) for _ in range(T)]
    par = list(ins_n)
    if rng.random() < 0.7:
        par[rng.randrange(T)] -= 1.0
    rot_n = [rng.random() * 6 for _ in range(T)]
    ctrl = [rng.random() * 20 for _ in range(2 * T)]
    r = rng.random()
    out, tube = get_single_tube_value(list(ctrl), ins_n, rot_n, par, 0.8, r)
    cases["single_tube"].append({"ctrl": ctrl, "ins": ins_n, "rot": rot_n, "parent": par, "r": r, "out": list(out), "tube": int(tube)})

HDICT = {"square_obstacle_avg_plus_weighted_goal": ["goal_weight"], "only_goal_distance": [],
         "follow_the_leader": ["only_tip"], "follow_the_leader_w_insertion": ["only_tip", "insertion_weight"],
         "follow_the_leader_translation": ["only_tip"]}
nrng = np.random.default_rng(7)
for name in HDICT:
    for only_tip in (True, False):
        conf = {"heuristic_type": name, "goal_weight": 3.0, "only_tip": only_tip, "insertion_weight": 0.25}
        f = create_heuristic_factory(conf, HDICT)
        # a chain root -> n1 -> n2 -> n3, a rewire test of n3 under n1, and create_from_old
        nodes, rec = [], []
        for d in range(4):
            ftl = [[nrng.normal(size=6) * (0 if (t + k + d) % 5 == 0 else 1) for k in range(3)] for t in range(2)]
            if d == 2:
                ftl[-1][-1][0:3] = 0.0  # pure translation screw
            kw = {"min_obstacle_distance": float(nrng.uniform(0.2, 5)), "goal_distance": float(nrng.uniform(0, 30)),
                  "follow_the_leader": ftl, "insertion_fraction": float(nrng.uniform(0, 2))}
            h = f.create(**kw)
            entry = {"kw": {k: (v if k != "follow_the_leader" else [[a.tolist() for a in t] for t in v]) for k, v in kw.items()},
                     "own_cost": h.get_cost()}
            if nodes:
                with redirect_stdout(io.StringIO()):
                    entry["test_cost"] = h.test_cost_from_parent(nodes[-1])
                    h.calculate_cost_from_parent(nodes[-1], init_insertion=(d == 3 and only_tip))
                entry["cost_after"] = h.get_cost()
                entry["generation"] = getattr(h, "generation", None)
            nodes.append(h)
            rec.append(entry)
        with redirect_stdout(io.StringIO()):
            new_ftl = [[nrng.normal(size=6) for k in range(2)] for t in range(2)]
            old = f.create_from_old(nodes[3], follow_the_leader=new_ftl)
            re_cost = old.test_cost_from_parent(nodes[1])
            old.calculate_cost_from_parent(nodes[1], reset=True)
        cases["heuristic"].append({"name": name, "only_tip": only_tip, "conf": conf, "chain": rec,
                                   "rewire_ftl": [[a.tolist() for a in t] for t in new_ftl], "rewire_test_cost": re_cost,
                                   "rewire_cost_after": old.get_cost()})

OUT.write_text(json.dumps(cases))
print("wrote", OUT, {k: len(v) for k, v in cases.items()})
