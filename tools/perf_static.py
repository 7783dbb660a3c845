"""Device timing of the static solve (+ fused static eval) for C1 (2 tubes) and C3 (3 tubes)."""
import ctypes as C
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "ctr-design-and-path-plan_b200"), str(ROOT / "tests")]
from util import gpu_static  # noqa: E402
from ctrdapp_b200.collision import CollisionChecker  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
for name, T, rad, env in (("c1_two_tube", 2, [0.9, 0.6], "env_boxes_multi.json"),
                          ("c3_three_tube", 3, [1.2, 0.9, 0.6], "env_c3_colon_straight.json")):
    m = gpu_static(name)
    N = np.asarray(m.num_discrete_points)
    rng = np.random.default_rng(1003)
    for lo, hi in ((1, 8), (1, 33)):
        e = np.stack([rng.integers(lo, min(hi, N[t] - 1) + 1, B) for t in range(T)], 1)
        idx = (N[None, :] - 1 - e).astype(np.int32)
        th = rng.uniform(0, 2 * np.pi, (B, T))
        m.solve_g_batch(idx[:1000], th[:1000], want_g=False)
        t0 = time.perf_counter()
        res = m.solve_g_batch(idx, th, want_g=False)
        dt = time.perf_counter() - t0
        cc = CollisionChecker(ROOT / "tests/fixtures" / env, ctx=m.ctx)
        t0 = time.perf_counter()
        obs, goal, cost, coll = cc.evaluate_batch(m, idx, th, rad, goal_weight=3.0)
        dt2 = time.perf_counter() - t0
        print(f"{name} exposure {lo}-{hi}: solve {B / dt / 1e3:.1f} kcfg/s  fused eval {B / dt2 / 1e3:.1f} kcfg/s  "
              f"converged {res['converged'].mean():.3f} mean nfev {res['nfev'].mean():.1f} collide {(obs < 0).mean():.3f}")
