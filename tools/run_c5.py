"""BASELINE configs[4]: Nelder-Mead design optimisation on colon_angled.STL -- every objective evaluation is a whole
batched RRT run for one design q; simplex vertices evaluated concurrently, one per rank (torchrun), RRT batches on each
rank's GPU.

    python tools/run_c5.py                                   # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_c5.py
"""
import json
import os
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "ctr-design-and-path-plan_b200")]
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

from ctrdapp_b200 import _lib as L  # noqa: E402
from ctrdapp_b200.collision import CollisionChecker  # noqa: E402
from ctrdapp_b200.heuristic import create_heuristic_factory  # noqa: E402
from ctrdapp_b200.optimize import create_optimizer  # noqa: E402

conf = {"optimizer_type": "nelder_mead", "optimizer_precision": 0.1, "optimize_iterations": int(os.environ.get("C5_EVALS", 60)),
        "solver_type": "rrt", "model_type": "kinematic", "heuristic_type": "square_obstacle_avg_plus_weighted_goal",
        "goal_weight": 3.0, "tube_number": 3, "tube_radius": {"outer": [1.2, 0.9, 0.6], "inner": [1.0, 0.7, 0.4]},
        "tube_lengths": [33, 66, 99], "delta_x": 1, "strain_bases": ["linear"] * 3, "q_dof": [2, 3, 3],
        "iteration_number": int(os.environ.get("C5_RRT", 2000)), "batch_size": 250, "step_bound": 7.5, "rotation_max": 0.5,
        "single_tube_control": False, "seed": 5}
HDICT = {"square_obstacle_avg_plus_weighted_goal": ["goal_weight"]}
ctx = L.Context(local)
cc = CollisionChecker(ROOT / "tests/fixtures/env_colon_angled.json", ctx=ctx)
guess = [0.02, 0.001, 0.05, 0.0001, 0.002, 0.03, 0.0002, 0.0005]
if os.environ.get("C5_SPECULATE") is not None:
    conf["speculate"] = int(os.environ["C5_SPECULATE"])
dev = torch.device("cuda", local)
# warm-up outside the timed region: one objective evaluation per rank (scratch buffers, kernel loading) and the first
# collective (NCCL builds its communicator lazily, ~1 s)
warm = create_optimizer(create_heuristic_factory(conf, HDICT), cc, guess, conf, device=dev)
warm.solver_heuristic(np.asarray(guess))
if world > 1:
    dist.all_reduce(torch.zeros(1, device=dev))
    dist.barrier()
torch.cuda.synchronize()
opt = create_optimizer(create_heuristic_factory(conf, HDICT), cc, guess, conf, device=dev)
t0 = time.perf_counter()
res = opt.find_min()
dt = time.perf_counter() - t0
if rank == 0:
    costs = [e["cost"] for e in res.optimize_process]           # what the sequential algorithm (SciPy's, the reference's) evaluates
    spec = [e["cost"] for e in res.speculative_results]         # evaluated on otherwise idle ranks and not asked for afterwards
    print(json.dumps({"workload": "c5_nelder_mead_colon_angled", "n_gpus": world, "seconds": dt,
                      "objective_evaluations": len(costs), "sequential_nfev": res.num_function_eval,
                      "speculative_discarded": len(spec), "iterations": res.num_iterations,
                      "rrt_iterations_per_objective": conf["iteration_number"],
                      "rrt_iterations_per_s_useful": len(costs) * conf["iteration_number"] / dt,
                      "rrt_iterations_per_s_incl_speculative": (len(costs) + len(spec)) * conf["iteration_number"] / dt,
                      "best_cost": float(min(costs)), "best_q": [float(v) for v in res.best_q],
                      "best_solver_rank": res.best_rank, "first_vertex_cost": float(costs[0])}))
if world > 1:
    dist.destroy_process_group()
