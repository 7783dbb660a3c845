"""Collision kernel on GIVEN curves (the path the static model takes) vs the fused kinematic path, same configurations:
separates "curves are read from HBM" from "static equilibrium shapes are harder".  usage: [B]"""
import ctypes as C
import pathlib
import sys

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "ctr-design-and-path-plan_b200"), str(ROOT / "tests")]
import bench  # noqa: E402
from ctrdapp_b200 import _lib as L  # noqa: E402
from ctrdapp_b200.collision import CollisionChecker  # noqa: E402
from ctrdapp_b200.model.kinematic import Kinematic  # noqa: E402
from util import gpu_static  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dev = torch.device("cuda", 0)
ms = gpu_static("c3_three_tube")
ctx = ms.ctx
m = Kinematic(3, [np.array(bench.Q[0:2]), np.array(bench.Q[2:5]), np.array(bench.Q[5:8])], bench.DOF, bench.LENGTHS,
              bench.DX, bench.BASES, ctx=ctx)
cc = CollisionChecker(bench.ENV_FILE, ctx=ctx)
lib = L.lib()
ih, th = bench.make_inputs(B, 1003, 1, 8)
idd, thd = torch.from_numpy(ih).to(dev), torch.from_numpy(th).to(dev)
rad = np.ascontiguousarray(bench.RAD, np.float64)
off = torch.zeros(B * 3 + 1, dtype=torch.int64, device=dev)
L.check(lib.ctrb_g_node_offsets(ctx.handle, m.handle, B, L.ptr(idd), 0, L.ptr(off), L.MEM_DEVICE))
total = int(off[-1].item())
g = torch.empty((total, 4, 4), dtype=torch.float64, device=dev)
obs = torch.empty(B, dtype=torch.float64, device=dev)
goal = torch.empty(B, dtype=torch.float64, device=dev)
cost = torch.empty(B, dtype=torch.float64, device=dev)
coll = torch.empty(B, dtype=torch.uint8, device=dev)
cd = L.CostDesc(0, 3.0, 1, 0.0)
# kinematic curves, given
L.check(lib.ctrb_solve_g_batch(ctx.handle, m.handle, B, L.ptr(idd), L.ptr(thd), 0, 64, L.ptr(off), L.ptr(g), L.MEM_DEVICE))
for _ in range(2):
    L.check(lib.ctrb_check_collision_batch(ctx.handle, cc.handle, B, 3, L.ptr(g), L.ptr(off), L.ptr(rad), L.ptr(obs), L.ptr(goal),
                                           L.MEM_DEVICE))
torch.cuda.synchronize()
w = ctx.work_counters()
print(f"kinematic curves from HBM : {ctx.last_kernel_ms('chain_eval'):7.2f} ms  collide {float((obs < 0).double().mean()):.3f}  "
      f"nodes {total / B:.1f}/cfg  bvh {w['bvh_nodes'] / B:.0f} pre {w['prefilter_tests'] / B:.1f} exact {w['exact_tests'] / B:.2f}")
for _ in range(2):
    L.check(lib.ctrb_eval_batch(ctx.handle, m.handle, cc.handle, C.byref(cd), B, L.ptr(idd), L.ptr(thd), L.ptr(rad), L.ptr(obs),
                                L.ptr(goal), L.ptr(cost), L.ptr(coll), L.MEM_DEVICE))
torch.cuda.synchronize()
print(f"kinematic fused           : {ctx.last_kernel_ms('chain_eval'):7.2f} ms  collide {float((obs < 0).double().mean()):.3f}")
# static curves, given
info = torch.empty(B, dtype=torch.int32, device=dev)
nfev = torch.empty(B, dtype=torch.int32, device=dev)
for _ in range(2):
    L.check(lib.ctrb_static_eval_batch(ctx.handle, ms.handle, cc.handle, C.byref(cd), B, L.ptr(idd), L.ptr(thd), L.ptr(rad),
                                       L.ptr(obs), L.ptr(goal), L.ptr(cost), L.ptr(coll), L.ptr(info), L.ptr(nfev), L.MEM_DEVICE))
torch.cuda.synchronize()
w = ctx.work_counters()
print(f"static curves from HBM    : {ctx.last_kernel_ms('chain_eval'):7.2f} ms  collide {float((obs < 0).double().mean()):.3f}  "
      f"bvh {w['bvh_nodes'] / B:.0f} pre {w['prefilter_tests'] / B:.1f} exact {w['exact_tests'] / B:.2f}")
