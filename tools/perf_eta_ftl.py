"""eta / follow-the-leader kernel alone: C2 model, 10^6 candidates, parents resident in HBM (what bench.py's extra reports)."""
import ctypes as C
import pathlib
import sys

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "ctr-design-and-path-plan_b200")]
from ctrdapp_b200 import _lib as L  # noqa: E402
from ctrdapp_b200.model.kinematic import Kinematic  # noqa: E402

dev = torch.device("cuda", 0)
ctx = L.Context(0)
lib = L.lib()
m2 = Kinematic(2, [np.array([0.02, 0.001]), np.array([0.05, 0.0001, 0.002])], [2, 3], [120, 120], 1, ["linear", "linear"], ctx=ctx)
rng = np.random.default_rng(1002)
B2 = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
prev_idx = torch.from_numpy(rng.integers(0, 120, (B2, 2)).astype(np.int32)).to(dev)
t2 = torch.from_numpy(rng.uniform(0, 2 * np.pi, (B2, 2))).to(dev)
off = torch.empty(B2 * 2 + 1, dtype=torch.int64, device=dev)
L.check(lib.ctrb_g_node_offsets(ctx.handle, m2.handle, B2, L.ptr(prev_idx), 0, L.ptr(off), L.MEM_DEVICE))
ctx.synchronize()
total = int(off[-1].item())
g = torch.empty((total, 4, 4), dtype=torch.float64, device=dev)
L.check(lib.ctrb_solve_g_batch(ctx.handle, m2.handle, B2, L.ptr(prev_idx), L.ptr(t2), 0, 64, L.ptr(off), L.ptr(g), L.MEM_DEVICE))
new_idx = (prev_idx - 1).clamp(min=0).contiguous()
dth = torch.full((B2, 2), 0.05, dtype=torch.float64, device=dev)
eta = torch.empty((B2, 2, 6), dtype=torch.float64, device=dev)
ftl = torch.empty((total, 6), dtype=torch.float64, device=dev)
mag = torch.empty((B2, 4), dtype=torch.float64, device=dev)
for want_ftl in (True, False):
    for rep in range(3):
        L.check(lib.ctrb_eta_ftl_batch(ctx.handle, m2.handle, B2, L.ptr(prev_idx), L.ptr(new_idx), None, L.ptr(dth), L.ptr(g),
                                       L.ptr(off), L.ptr(eta), L.ptr(ftl) if want_ftl else None, L.ptr(mag), L.MEM_DEVICE))
        ms = ctx.last_kernel_ms("eta_ftl")
    alg = total * (96 + (48 if want_ftl else 0))
    print(f"ftl_out={want_ftl}: {ms:.3f} ms, {total} parent nodes, {alg / ms / 1e6:.0f} GB/s algorithmic ({B2 / ms / 1e3:.1f} M candidates/s)", flush=True)
assert ctx.device_status() == 0
print("checksum", float(eta.sum()), float(mag.sum()))
