"""One fused kinematic evaluation of B C3 configurations (for ncu captures of the collision kernel).
usage: python tools/ncu_eval.py [B] [lo] [hi]"""
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

B = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1, 8)
dev = torch.device("cuda", 0)
ctx = L.Context(0)
m = Kinematic(3, [np.array(bench.Q[0:2]), np.array(bench.Q[2:5]), np.array(bench.Q[5:8])], bench.DOF, bench.LENGTHS,
              bench.DX, bench.BASES, ctx=ctx)
cc = CollisionChecker(bench.ENV_FILE, ctx=ctx)
rad = np.ascontiguousarray(bench.RAD, np.float64)
cd = L.CostDesc(0, 3.0, 1, 0.0)
out = [torch.empty(B, dtype=torch.float64, device=dev) for _ in range(3)] + [torch.empty(B, dtype=torch.uint8, device=dev)]
ih, th = bench.make_inputs(B, 1003, lo, hi)
idd, thd = torch.from_numpy(ih).to(dev), torch.from_numpy(th).to(dev)
for rep in range(3):
    L.check(L.lib().ctrb_eval_batch(ctx.handle, m.handle, cc.handle, C.byref(cd), B, L.ptr(idd), L.ptr(thd),
                                    L.ptr(rad), L.ptr(out[0]), L.ptr(out[1]), L.ptr(out[2]), L.ptr(out[3]), 1))
torch.cuda.synchronize()
print("collide", float((out[0] < 0).double().mean()), ctx.work_counters())
