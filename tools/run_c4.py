"""BASELINE configs[3]: RRT* with batched candidate extension and the follow-the-leader heuristic against
follow_the_leader_goal_multiple.STL -- the reference's test/configuration/config_ftl.yaml (1 tube of radius 5, linear
strain base with 2 dof, delta_x 2.5, insertion_max 50, step_bound 7.5, 4000 iterations, rrt_star,
follow_the_leader_translation, only_tip; `tube_lengths: 50` added, the file fails validation as shipped,
parse_config.py:152-155), goal mesh from init_objects_ftl_goal.json, no obstacles.  The reference's RNG is unseeded
(rrt.py:5): the comparison is statistical -- iterations per second, nodes kept, whether the goal mesh was reached, and
the cost of the best path -- for several batch sizes, next to the same search driven one candidate at a time through the
CPU oracle (the reference's call pattern).

    python tools/run_c4.py [q0 q1]
"""
import json
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "ctr-design-and-path-plan_b200"), str(ROOT / "tests")]
from ctrdapp_b200.collision import CollisionChecker  # noqa: E402
from ctrdapp_b200.heuristic import create_heuristic_factory  # noqa: E402
from ctrdapp_b200.model.kinematic import Kinematic  # noqa: E402
from ctrdapp_b200.solve import create_solver  # noqa: E402

q = [float(v) for v in sys.argv[1:3]] or [0.01, 0.0005]
CONF = {"solver_type": "rrt_star", "model_type": "kinematic", "heuristic_type": "follow_the_leader_translation",
        "tube_number": 1, "tube_radius": {"outer": [5.0], "inner": [4.0]}, "tube_lengths": [50], "delta_x": 2.5,
        "strain_bases": ["linear"], "q_dof": [2], "step_bound": 7.5, "rotation_max": 0.1745, "insertion_max": 50,
        "iteration_number": 4000, "single_tube_control": False, "rewire_probability": 0.1, "only_tip": True,
        "insertion_weight": 10, "goal_weight": 3, "seed": 4}
HDICT = {"follow_the_leader_translation": ["only_tip"]}
env = ROOT / "tests/fixtures/env_ftl_goal.json"
m = Kinematic(1, [np.asarray(q)], [2], [50], 2.5, ["linear"])
cc = CollisionChecker(env, ctx=m.ctx)
fac = create_heuristic_factory(CONF, HDICT)
rows = []
for K in (1, 16, 64, 256, 1000):
    conf = dict(CONF, batch_size=K)
    create_solver(m, fac, cc, dict(conf, iteration_number=min(4000, 4 * K)))  # warm-up
    t0 = time.perf_counter()
    s = create_solver(m, fac, cc, conf)
    dt = time.perf_counter() - t0
    best = s.get_best_cost()
    best = float(best[0] if isinstance(best, (tuple, list)) else best)
    rows.append({"batch_size": K, "seconds": dt, "iterations_per_s": CONF["iteration_number"] / dt,
                 "nodes_kept": len(s.tree.nodes), "goal_reached": bool(s.found_solution), "best_cost": best})
    print(json.dumps(rows[-1]), flush=True)

# the same search, one candidate at a time through the CPU oracle (test infrastructure; the reference's call pattern)
from oracle import oracle as orc  # noqa: E402


try:
    from test_gpu_solve import OracleModel  # noqa: E402
    import models_def  # noqa: E402
    models_def.MODELS["c4_one_tube"] = (["linear"], [2], q, [50], 2.5)
    om = OracleModel("c4_one_tube")
    conf = dict(CONF, batch_size=1, iteration_number=1000)
    t0 = time.perf_counter()
    s = create_solver(om, fac, orc.Env.from_json(env), conf)
    dt = time.perf_counter() - t0
    print(json.dumps({"cpu_oracle_one_candidate_at_a_time": True, "iterations_per_s": 1000 / dt, "nodes_kept": len(s.tree.nodes),
                      "goal_reached": bool(s.found_solution)}), flush=True)
except Exception as e:  # the oracle arm is optional colour, the GPU arm above is the workload
    print(json.dumps({"cpu_oracle_arm": "skipped", "why": repr(e)[:200]}))
