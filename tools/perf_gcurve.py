"""Device-resident timing of the C2 g-curve kernels (BASELINE configs[1]: 2-tube kinematic model, 10^6 configurations,
fp64 and fp32 outputs) and a checksum of the output.  usage: [CTRB200_LIB=path] python tools/perf_gcurve.py [B] [full]"""
import pathlib
import sys

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "ctr-design-and-path-plan_b200")]
from ctrdapp_b200 import _lib as L  # noqa: E402
from ctrdapp_b200.model.kinematic import Kinematic  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
full = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = torch.device("cuda", 0)
ctx = L.Context(0)
lib = L.lib()
m2 = Kinematic(2, [np.array([0.02, 0.001]), np.array([0.05, 0.0001, 0.002])], [2, 3], [120, 120], 1, ["linear", "linear"], ctx=ctx)
rng = np.random.default_rng(1002)
i2 = torch.from_numpy(rng.integers(0, 121, (B, 2)).astype(np.int32)).to(dev)
t2 = torch.from_numpy(rng.uniform(0, 2 * np.pi, (B, 2))).to(dev)
off = torch.empty(B * 2 + 1, dtype=torch.int64, device=dev)
L.check(lib.ctrb_g_node_offsets(ctx.handle, m2.handle, B, L.ptr(i2), full, L.ptr(off), L.MEM_DEVICE))
total = int(off[-1].item())
for prec, dt_ in ((64, torch.float64), (32, torch.float32)):
    g = torch.empty((total, 4, 4), dtype=dt_, device=dev)
    ts = []
    for rep in range(6):
        ctx.mark(0)
        L.check(lib.ctrb_solve_g_batch(ctx.handle, m2.handle, B, L.ptr(i2), L.ptr(t2), full, prec, L.ptr(off), L.ptr(g), L.MEM_DEVICE))
        ctx.mark(1)
        ts.append(ctx.elapsed_ms())
    ms = float(np.median(ts[1:]))
    gbs = g.numel() * g.element_size() / (ms * 1e-3) / 1e9
    print(f"fp{prec}: {ms:.3f} ms (min {min(ts):.3f}) {B / ms / 1e3:.1f} Mcfg/s {gbs:.0f} GB/s nodes {total} "
          f"checksum {float(g.double().sum()):.12e} abs {float(g.double().abs().sum()):.12e}", flush=True)
    del g
