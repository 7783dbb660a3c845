"""Rank process of test_optimize_host.py::test_distributed_evaluator_two_ranks (gloo, CPU)."""
import pathlib
import sys

import numpy as np
import torch.distributed as dist

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "ctr-design-and-path-plan_b200"), str(ROOT / "tests")]
from ctrdapp_b200.optimize.simplex import DistributedEvaluator, nelder_mead  # noqa: E402
from test_optimize_host import rosen  # noqa: E402

rank, world, out = int(sys.argv[1]), int(sys.argv[2]), pathlib.Path(sys.argv[3])
dist.init_process_group("gloo", rank=rank, world_size=world)
ev = DistributedEvaluator(rosen)
simplex = np.random.default_rng(11).uniform(-1.5, 1.5, (6, 5))
res = nelder_mead(ev, simplex, xatol=1e-6, fatol=1e-6, maxfev=3000)
# the rank that holds the lowest local value (here: a made-up one per rank) is the same on both ranks
owner = ev.owner_of_lowest(3.0 - rank)
assert owner == world - 1 and ev.world_size() == world and ev.rank() == rank
np.savez(out / f"rank{rank}.npz", x=res.x, nit=res.nit, nfev=res.nfev, calls=ev.calls)
dist.destroy_process_group()
