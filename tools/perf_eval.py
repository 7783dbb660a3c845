"""Quick device-resident timing of the fused kernel for both C3 exposure distributions.
usage: [CTRB200_LIB=path] python tools/perf_eval.py [B]"""
import ctypes as C
import pathlib
import sys

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "ctr-design-and-path-plan_b200")]
import bench  # noqa: E402
from ctrdapp_b200 import _lib as L  # noqa: E402
from ctrdapp_b200.collision import CollisionChecker  # noqa: E402
from ctrdapp_b200.model.kinematic import Kinematic  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
dev = torch.device("cuda", 0)
ctx = L.Context(0)
m = Kinematic(3, [np.array(bench.Q[0:2]), np.array(bench.Q[2:5]), np.array(bench.Q[5:8])], bench.DOF, bench.LENGTHS,
              bench.DX, bench.BASES, ctx=ctx)
envs = {"c3": bench.ENV_FILE, "new": ROOT / "tests/fixtures/env_colon_straight_new.json"}
rad = np.ascontiguousarray(bench.RAD, np.float64)
cd = L.CostDesc(0, 3.0, 1, 0.0)
out = [torch.empty(B, dtype=torch.float64, device=dev) for _ in range(3)] + [torch.empty(B, dtype=torch.uint8, device=dev)]
res = []
for ename, efile in envs.items():
    cc = CollisionChecker(efile, ctx=ctx)
    for wname, (lo, hi) in (("shallow", (1, 8)), ("uniform", (1, 33))):
        ih, th = bench.make_inputs(B, 1003, lo, hi)
        idd, thd = torch.from_numpy(ih).to(dev), torch.from_numpy(th).to(dev)
        best = 1e30
        for rep in range(3):
            ctx.mark(0)
            L.check(L.lib().ctrb_eval_batch(ctx.handle, m.handle, cc.handle, C.byref(cd), B, L.ptr(idd), L.ptr(thd),
                                            L.ptr(rad), L.ptr(out[0]), L.ptr(out[1]), L.ptr(out[2]), L.ptr(out[3]), 1))
            ctx.mark(1)
            best = min(best, ctx.elapsed_ms())
        w = ctx.work_counters()
        res.append(f"{ename}/{wname}: {B / best / 1e3:8.2f} Mcfg/s  collide {float((out[0] < 0).double().mean()):.3f} "
                   f"nodes/cfg {w['bvh_nodes'] / B:.0f} pre {w['prefilter_tests'] / B:.1f} exact {w['exact_tests'] / B:.2f} "
                   f"chk {float(out[0][out[0] > 0].sum()):.6f}")
print(L.lib_path().name, cc.stats())
print("\n".join(res))
