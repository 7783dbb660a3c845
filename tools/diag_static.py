import sys, pathlib
import numpy as np
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "ctr-design-and-path-plan_b200"), str(ROOT / "tests")]
from util import gpu_static, orc_static
from oracle import oracle as orc
name = "c1_two_tube"
m, s = gpu_static(name), orc_static(name)
rng = np.random.default_rng(1001)
B = 300
idx = rng.integers(0, 118, (B, 2)).astype(np.int32)
th = rng.uniform(0, 2 * np.pi, (B, 2))
res = m.solve_g_batch(idx, th)
x0 = m.general_init_guess
F0 = np.linalg.norm(orc.static_residual(s, idx[0], th[0], x0))
n = 0
for b in range(B):
    g, sol, nfev, info = orc.static_solve_g(s, idx[b], th[b])
    if info == 1 and res["converged"][b]:
        d = np.max(np.abs(sol - res["solution"][b])) / max(1.0, np.max(np.abs(sol)))
        if d > 1e-5:
            Fo = np.linalg.norm(orc.static_residual(s, idx[b], th[b], sol))
            Fg = np.linalg.norm(orc.static_residual(s, idx[b], th[b], res["solution"][b]))
            Fg2 = np.linalg.norm(m.static_equilibrium_batch([idx[b]], [th[b]], [res["solution"][b]])[0])
            print(b, idx[b].tolist(), "d", d, "|F| oracle-root", Fo, "gpu-root", Fg, Fg2, "nfev", nfev, res["nfev"][b],
                  "max|sol|", np.max(np.abs(sol)), np.max(np.abs(res["solution"][b])))
            n += 1
            if n < 3:
                print("  orc", np.array2string(sol, precision=5))
                print("  gpu", np.array2string(res["solution"][b], precision=5))
print("mismatch", n, "F at x0", F0)
