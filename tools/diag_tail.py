"""Tail diagnosis of the static solve: per seed (= per rank of bench.py) the kernel times, the fallback count and the
largest evaluation counts.  usage: diag_tail.py [B] [first seed] [number of seeds]"""
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "ctr-design-and-path-plan_b200"), str(ROOT / "tests")]
from util import gpu_static  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1003
nseeds = int(sys.argv[3]) if len(sys.argv) > 3 else 8
m = gpu_static("c3_three_tube")
N = np.asarray(m.num_discrete_points)
for seed in range(seed0, seed0 + nseeds):
    rng = np.random.default_rng(seed)
    e = np.stack([rng.integers(1, min(8, N[t] - 1) + 1, B) for t in range(3)], 1)
    idx = (N[None, :] - 1 - e).astype(np.int32)
    th = rng.uniform(0, 2 * np.pi, (B, 3))
    m.solve_g_batch(idx[:4096], th[:4096], want_g=False)
    res = m.solve_g_batch(idx, th, want_g=False)
    tot, fast, fb = m.ctx.last_kernel_ms("static_solve"), m.ctx.last_kernel_ms("static_fast"), m.ctx.static_fallbacks()
    nf = res["nfev"]
    top = np.argsort(nf)[::-1][:6]
    print(f"seed {seed}: solve {tot:.1f} ms (fast {fast:.1f}, QR {tot - fast:.1f}) fallbacks {fb} mean nfev {nf.mean():.1f} "
          f"info {np.bincount(res['info'], minlength=6).tolist()} top nfev {nf[top].tolist()} at {top.tolist()} "
          f"idx {idx[top[0]].tolist()} >1000: {(nf > 1000).sum()}", flush=True)
