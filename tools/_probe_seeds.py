import ctypes as C, pathlib, sys
import numpy as np, torch
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "ctr-design-and-path-plan_b200"), str(ROOT / "tests")]
import bench
from ctrdapp_b200 import _lib as L
from ctrdapp_b200.collision import CollisionChecker
from util import gpu_static
B = 1 << 20
dev = torch.device("cuda", 0)
ms = gpu_static("c3_three_tube")
ctx = ms.ctx
cc = CollisionChecker(bench.ENV_FILE, ctx=ctx)
lib = L.lib()
rad = np.ascontiguousarray(bench.RAD, np.float64)
cd = L.CostDesc(0, 3.0, 1, 0.0)
obs = torch.empty(B, dtype=torch.float64, device=dev); goal = torch.empty_like(obs); cost = torch.empty_like(obs)
coll = torch.empty(B, dtype=torch.uint8, device=dev); info = torch.empty(B, dtype=torch.int32, device=dev); nfev = torch.empty_like(info)
for seed in [int(a) for a in sys.argv[1:]] or [1003 + r for r in range(8)]:
    ih, th = bench.make_inputs(B, seed, 1, 8)
    idd, thd = torch.from_numpy(ih).to(dev), torch.from_numpy(th).to(dev)
    for _ in range(2):
        L.check(lib.ctrb_static_eval_batch(ctx.handle, ms.handle, cc.handle, C.byref(cd), B, L.ptr(idd), L.ptr(thd), L.ptr(rad),
                                           L.ptr(obs), L.ptr(goal), L.ptr(cost), L.ptr(coll), L.ptr(info), L.ptr(nfev), L.MEM_DEVICE))
    torch.cuda.synchronize()
    w = ctx.work_counters()
    print(f"seed {seed}: solve {ctx.last_kernel_ms('static_solve'):7.2f} ms (fast {ctx.last_kernel_ms('static_fast'):6.2f}) collision {ctx.last_kernel_ms('chain_eval'):6.2f} ms "
          f"nfev max {int(nfev.max())} exact/cfg {w['exact_tests'] / B:.2f} pre/cfg {w['prefilter_tests'] / B:.1f}", flush=True)
