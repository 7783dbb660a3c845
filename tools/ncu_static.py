"""Small static-solve run for ncu: C3 shallow, B configs, device-side only."""
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "ctr-design-and-path-plan_b200"), str(ROOT / "tests")]
from util import gpu_static  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
LO = int(sys.argv[2]) if len(sys.argv) > 2 else 1
HI = int(sys.argv[3]) if len(sys.argv) > 3 else 8
m = gpu_static("c3_three_tube")
N = np.asarray(m.num_discrete_points)
rng = np.random.default_rng(1003)
e = np.stack([rng.integers(LO, HI + 1, B) for t in range(3)], 1)
idx = (N[None, :] - 1 - e).astype(np.int32)
th = rng.uniform(0, 2 * np.pi, (B, 3))
for _ in range(2):
    res = m.solve_g_batch(idx, th, want_g=True)
print("converged", res["converged"].mean(), "nfev", res["nfev"].mean(), "ms", m.ctx.last_kernel_ms("static_solve"),
      "fast ms", m.ctx.last_kernel_ms("static_fast"), "fallbacks", m.ctx.static_fallbacks())

# the fused static evaluation (solve -> curves -> collision -> cost): launches static_fast_kernel, static_solve_kernel,
# static_curves_kernel and chain_eval_kernel<false> once more -- the four launches an ncu capture with -s 6 -c 4 records
from ctrdapp_b200.collision import CollisionChecker  # noqa: E402
cc = CollisionChecker(ROOT / "tests/fixtures/env_c3_colon_straight.json", ctx=m.ctx)
obs, goal, cost, coll = cc.evaluate_batch(m, idx, th, [1.2, 0.9, 0.6], goal_weight=3.0)
print("collide", (obs < 0).mean(), "chain_eval ms", m.ctx.last_kernel_ms("chain_eval"), "work", m.ctx.work_counters())
