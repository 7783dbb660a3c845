"""Planner throughput: RRT iterations per second through the batched hot path for several batch sizes, next to the same
search driven one candidate at a time through the CPU oracle (the reference's call pattern)."""
import pathlib
import sys
import time

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "ctr-design-and-path-plan_b200"), str(ROOT / "tests")]
from test_gpu_solve import CONF, HDICT, OracleModel  # noqa: E402
from util import gpu_model  # noqa: E402
from ctrdapp_b200.collision import CollisionChecker  # noqa: E402
from ctrdapp_b200.heuristic import create_heuristic_factory  # noqa: E402
from ctrdapp_b200.solve import create_solver  # noqa: E402
from oracle import oracle as orc  # noqa: E402

env = ROOT / "tests/fixtures/env_c3_colon_straight.json"
fac = lambda name: create_heuristic_factory(dict(CONF, heuristic_type=name), HDICT)  # noqa: E731
m, cc = gpu_model("c3_three_tube"), CollisionChecker(env)
for solver, heur in (("rrt", "square_obstacle_avg_plus_weighted_goal"), ("rrt_star", "follow_the_leader")):
    for K, iters in ((1, 2000), (64, 8192), (512, 16384), (4096, 32768)):
        conf = dict(CONF, solver_type=solver, batch_size=K, iteration_number=iters)
        create_solver(m, fac(heur), cc, dict(conf, iteration_number=min(iters, 4 * K)))  # warm-up
        t0 = time.perf_counter()
        s = create_solver(m, fac(heur), cc, conf)
        dt = time.perf_counter() - t0
        print(f"{solver:8s} {heur[:18]:18s} batch {K:5d}: {iters / dt:10.0f} iterations/s  ({len(s.tree.nodes)} nodes kept, "
              f"goal reached {s.found_solution})", flush=True)
    conf = dict(CONF, solver_type=solver, batch_size=1, iteration_number=1000)
    t0 = time.perf_counter()
    s = create_solver(OracleModel("c3_three_tube"), fac(heur), orc.Env.from_json(env), conf)
    dt = time.perf_counter() - t0
    print(f"{solver:8s} {heur[:18]:18s} CPU oracle, one candidate at a time: {1000 / dt:10.0f} iterations/s ({len(s.tree.nodes)} nodes)",
          flush=True)
