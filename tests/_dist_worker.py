"""Worker of tests/test_dist_cpu.py: one rank of a gloo group (launched like torchrun launches bench.py)."""
import json
import os
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "ctr-design-and-path-plan_b200"), str(ROOT)]

import torch.distributed as dist  # noqa: E402
from ctrdapp_b200 import dist as cdist  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    costs = np.load(sys.argv[1])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = cdist.shard_range(len(costs), rank, world)
        c, i = cdist.local_best(costs[lo:hi], lo)
        best = cdist.gather_best(c, i)
        # the same through the device-shaped path: this rank's [key, index] tensor, two all-reduce(min) steps
        import torch
        t = torch.tensor([cdist.cost_to_key(c) if i >= 0 else cdist.INT64_MAX, i if i >= 0 else cdist.INT64_MAX], dtype=torch.int64)
        assert cdist.decode_best(cdist.reduce_best(t)) == best
        per = len(costs) // world
        allc = cdist.gather_costs(costs[rank * per:(rank + 1) * per]).numpy()
        print(json.dumps({"rank": rank, "best": best, "costs_ok": bool(np.array_equal(allc, costs[:per * world], equal_nan=True))}),
              flush=True)
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
