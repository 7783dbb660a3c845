"""Device timing (CUDA events inside the library) of the static solve: explicit-inverse kernel, QR kernel, fallbacks."""
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "ctr-design-and-path-plan_b200"), str(ROOT / "tests")]
from util import gpu_static  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
import os
WANT_G = bool(int(os.environ.get("WANT_G", "0")))
cases = [a for a in sys.argv[2:]] or ["c1:1:8", "c1:1:33", "c3:1:8", "c3:2:8", "c3:1:33", "c3:2:33"]
models = {"c1": gpu_static("c1_two_tube"), "c3": gpu_static("c3_three_tube")}
for case in cases:
    name, lo, hi = case.split(":")
    lo, hi = int(lo), int(hi)
    m = models[name]
    T = m.tube_num
    N = np.asarray(m.num_discrete_points)
    rng = np.random.default_rng(1003)
    e = np.stack([rng.integers(lo, min(hi, N[t] - 1) + 1, B) for t in range(T)], 1)
    idx = (N[None, :] - 1 - e).astype(np.int32)
    th = rng.uniform(0, 2 * np.pi, (B, T))
    m.solve_g_batch(idx[:2000], th[:2000], want_g=False)
    res = m.solve_g_batch(idx, th, want_g=WANT_G)
    tot, fast, fb = m.ctx.last_kernel_ms("static_solve"), m.ctx.last_kernel_ms("static_fast"), m.ctx.static_fallbacks()
    print(f"{case}: total {tot:.1f} ms = {B / tot / 1e3:.2f} Mcfg/s | fast kernel {fast:.1f} ms "
          f"({(B - fb) / max(fast, 1e-9) / 1e3:.2f} Mcfg/s over {B - fb}) | QR kernel {tot - fast:.1f} ms "
          f"({fb / max(tot - fast, 1e-9) / 1e3:.2f} Mcfg/s over {fb} = {100 * fb / B:.1f} %) | "
          f"converged {res['converged'].mean():.3f} nfev {res['nfev'].mean():.1f}")
