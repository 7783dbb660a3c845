"""Device timing of the static solve per kernel (CUDA events inside the library) for both treatments of one-step
sections, with the distribution of the evaluation counts.  usage: perf_static3.py [B] [case ...]   case = lo:hi"""
import hashlib
import json
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "ctr-design-and-path-plan_b200"), str(ROOT / "tests")]
from util import gpu_static  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
cases = sys.argv[2:] or ["1:8", "1:33", "2:8"]
m = gpu_static("c3_three_tube")
N = np.asarray(m.num_discrete_points)
for case in cases:
    lo, hi = (int(v) for v in case.split(":"))
    rng = np.random.default_rng(1003)
    e = np.stack([rng.integers(lo, min(hi, N[t] - 1) + 1, B) for t in range(3)], 1)
    idx = (N[None, :] - 1 - e).astype(np.int32)
    th = rng.uniform(0, 2 * np.pi, (B, 3))
    short = np.any(e[:, :2] == 1, axis=1)
    for one_step in (0, 1):
        with m.ctx.options(static_one_step=one_step):
            m.solve_g_batch(idx[:4096], th[:4096], want_g=False)
            res = m.solve_g_batch(idx, th, want_g=False)
            k = {n: m.ctx.last_kernel_ms(n) for n in ("static_solve", "static_fast", "static_qr")}
            fb = m.ctx.static_fallbacks()
        nf = res["nfev"]
        out = {"case": case, "B": B, "one_step_option": one_step, "ms": {a: round(b, 2) for a, b in k.items()},
               "Mcfg_per_s": round(B / k["static_solve"] / 1e3, 3), "fallbacks": fb, "one_step_fraction": round(float(short.mean()), 4),
               "converged": round(float(res["converged"].mean()), 4),
               "nfev": {"mean": round(float(nf.mean()), 1), "p50": int(np.median(nf)), "p99": int(np.quantile(nf, 0.99)), "max": int(nf.max()),
                        "mean_one_step": round(float(nf[short].mean()), 1) if short.any() else None,
                        "mean_other": round(float(nf[~short].mean()), 1)},
               "info_hist": np.bincount(res["info"], minlength=6).tolist(), "fallback_reasons": m.ctx.static_fallback_reasons(),
               # bit-level fingerprint of the whole batch: kernel variants that only re-schedule the same arithmetic must agree
               "md5": hashlib.md5(res["solution"].tobytes() + nf.tobytes() + res["info"].tobytes()).hexdigest()[:12]}
        print(json.dumps(out), flush=True)
