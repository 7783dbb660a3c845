"""Per-kernel device time of the fused static call (solve -> curves -> collision -> cost) on bench.py's inputs.
usage: perf_static_eval.py [B] [lo:hi ...]"""
import ctypes as C
import json
import pathlib
import sys

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "ctr-design-and-path-plan_b200"), str(ROOT / "tests")]
import bench  # noqa: E402
from ctrdapp_b200 import _lib as L  # noqa: E402
from ctrdapp_b200.collision import CollisionChecker  # noqa: E402
from ctrdapp_b200.model.static import Static  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
cases = sys.argv[2:] or ["1:8"]
dev = torch.device("cuda", 0)
ctx = L.Context(0)
qs = [np.array(bench.Q[0:2]), np.array(bench.Q[2:5]), np.array(bench.Q[5:8])]
ms = Static(3, qs, bench.DOF, bench.LENGTHS, bench.DX, bench.BASES, {"outer": bench.RAD, "inner": bench.RIN}, 2, ctx=ctx)
cc = CollisionChecker(bench.ENV_FILE, ctx=ctx)
cd = L.CostDesc(L.COST_TYPES["square_obstacle_avg_plus_weighted_goal"], bench.GOAL_WEIGHT, 1, 0.0)
rad = np.ascontiguousarray(bench.RAD, np.float64)
for case in cases:
    lo, hi = (int(v) for v in case.split(":"))
    idx, th = bench.make_inputs(B, 1003, lo, hi)
    idd, thd = torch.from_numpy(idx).to(dev), torch.from_numpy(th).to(dev)
    outs = [torch.empty(B, dtype=torch.float64, device=dev) for _ in range(3)] + [torch.empty(B, dtype=torch.uint8, device=dev)] \
        + [torch.empty(B, dtype=torch.int32, device=dev) for _ in range(2)]
    for rep in range(3):
        ctx.mark(0)
        L.check(L.lib().ctrb_static_eval_batch(ctx.handle, ms.handle, cc.handle, C.byref(cd), B, L.ptr(idd), L.ptr(thd), L.ptr(rad),
                                               *[L.ptr(o) for o in outs], L.MEM_DEVICE))
        ctx.mark(1)
        total = ctx.elapsed_ms()
    k = {n: round(ctx.last_kernel_ms(n), 3) for n in ("static_fast", "static_qr", "static_curves", "chain_eval")}
    print(json.dumps({"case": case, "B": B, "total_ms": round(total, 3), "Mcfg_per_s": round(B / total / 1e3, 3), "kernel_ms": k,
                      "collide": round(float((outs[0] < 0).double().mean()), 4), "work": ctx.work_counters(),
                      "checksum": float(torch.nan_to_num(outs[0]).sum())}), flush=True)
