"""Which headline configurations break the collision contract on the GPU's own unknowns, and how."""
import pathlib, sys, json
import numpy as np
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "ctr-design-and-path-plan_b200"), str(ROOT / "tests")]
import bench, static_parity as sp
from oracle import oracle as orc
from ctrdapp_b200.collision import CollisionChecker
from ctrdapp_b200.model.static import Static
RAD, RIN = [1.2, 0.9, 0.6], [1.0, 0.7, 0.4]
qs = [np.array(bench.Q[0:2]), np.array(bench.Q[2:5]), np.array(bench.Q[5:8])]
ms = Static(3, qs, bench.DOF, bench.LENGTHS, bench.DX, bench.BASES, {"outer": RAD, "inner": RIN}, 2)
cc = CollisionChecker(bench.ENV_FILE)
s = orc.make_static(bench.BASES, bench.DOF, bench.Q, bench.LENGTHS, bench.DX, RAD, RIN, 2)
oenv = orc.Env.from_json(bench.ENV_FILE)
for lo, hi, B in ((1, 8, 50000), (1, 33, 12000)):
    idx, theta = bench.make_inputs(B, 1003, lo, hi)
    obs, goal, cost, coll = cc.evaluate_batch(ms, idx, theta, RAD, goal_weight=3.0)
    res = ms.solve_g_batch(idx, theta, want_g=True)
    q = res["solution"]
    o = orc.static_batch(s, idx, theta, env=oenv, rad=RAD, goal_weight=3.0, q_in=q)
    bad, nb = sp.collision_contract(obs, goal, cost, o["obs"], o["goal"], o["cost"])
    print(lo, hi, "violations", int(bad.sum()), "in band", nb)
    N = np.asarray(bench.NPTS)
    for b in np.flatnonzero(bad)[:12]:
        T = 3
        l0, h0 = res["offsets"][b * T], res["offsets"][(b + 1) * T]
        g = res["g"][l0:h0]
        ref = np.concatenate(orc.static_curves(s, idx[b], theta[b], q[b]), 0)
        seg = np.linalg.norm(np.diff(g[:, 0:3, 3], axis=0), axis=1)
        print(json.dumps({"b": int(b), "e": (N - 1 - idx[b]).tolist(), "info": int(res["info"][b]), "qmax": float(np.abs(q[b]).max()),
                          "gpu": [float(obs[b]), float(goal[b]), float(cost[b])], "orc": [float(o["obs"][b]), float(o["goal"][b]), float(o["cost"][b])],
                          "curve_abs_err": float(np.abs(g - ref).max()), "pmax": float(np.abs(g[:, 0:3, 3]).max()), "max_seg": float(seg.max())}))
