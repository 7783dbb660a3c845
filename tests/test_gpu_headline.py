"""GPU parity on the HEADLINE workload: bench.py's own inputs (BASELINE configs[2]: 3-tube static model x
colon_straight.STL), through ctrb_static_eval_batch / ctrb_static_solve_g_batch, against the CPU oracle.

Reference path: Static.solve_g (static.py:71-135) -> CollisionChecker.check_collision
(collision_checker.py:35-70) -> own cost (obstacles_and_goal.py:39-45).

What is asserted (tests/static_parity.py has the bucket definitions):
  1. kernels, EVERY configuration: (a) the oracle queries the mesh with the curves the GPU integrated: collision bit
     exact outside the 1e-6 mm band, distances within 1e-6 mm, cost 1e-9 relative; (b) the oracle integrates the
     curves of the unknowns the GPU found: 1e-10 relative, widened only by the conditioning of integrate_g's
     absolute-index polynomials where one-step sections produce huge cancelling coefficients.
  2. solver, bucket A (same root as the oracle's hybrd): every configuration outside one-step sections, and every
     one-step configuration against the oracle's reduced-system hybrd, is in bucket A except for a fraction bounded
     by the oracle's OWN sensitivity, measured here by re-running it from an initial guess perturbed by 1e-13.
  3. solver, bucket B (another root): the GPU's unknowns are a root under the oracle's marched residual -- |F| within
     the range the oracle's own converged solves reach.
  4. one-step sections against the reference's literal hybrd: its null-space component is decided by rounding (its
     self-perturbation rate is measured and printed), so the statement is 3. plus the converged flag.
"""
import json
import os

import numpy as np
import pytest

import static_parity as sp
from conftest import FIXTURES, ROOT

pytestmark = pytest.mark.gpu

RAD, RIN, GOAL_WEIGHT = [1.2, 0.9, 0.6], [1.0, 0.7, 0.4], 3.0


def _objects():
    import bench
    from oracle import oracle as orc
    from ctrdapp_b200.collision import CollisionChecker
    from ctrdapp_b200.model.static import Static
    qs = [np.array(bench.Q[0:2]), np.array(bench.Q[2:5]), np.array(bench.Q[5:8])]
    ms = Static(3, qs, bench.DOF, bench.LENGTHS, bench.DX, bench.BASES, {"outer": RAD, "inner": RIN}, 2)
    cc = CollisionChecker(bench.ENV_FILE)
    s = orc.make_static(bench.BASES, bench.DOF, bench.Q, bench.LENGTHS, bench.DX, RAD, RIN, 2)
    oenv = orc.Env.from_json(bench.ENV_FILE)
    return bench, orc, ms, cc, s, oenv


@pytest.mark.parametrize("lo,hi,B", [(1, 8, 50_000), (1, 33, 12_000)])
def test_headline_static_parity(lo, hi, B):
    B = int(os.environ.get("CTRB_HEADLINE_PARITY_B", B))  # a larger one-off run for profiles/ (the oracle needs ~0.3 ms per configuration)
    bench, orc, ms, cc, s, oenv = _objects()
    idx, theta = bench.make_inputs(B, 1003, lo, hi)        # bench.py's exact inputs (rank 0)
    N = np.asarray(bench.NPTS)
    short = sp.one_step(idx, N)

    # ---- the product path
    obs, goal, cost, coll = cc.evaluate_batch(ms, idx, theta, RAD, goal_weight=GOAL_WEIGHT)
    res = ms.solve_g_batch(idx, theta, want_g=False)
    assert np.array_equal(res["info"], cc.last_static["info"])       # the fused call runs the same solve
    assert np.array_equal(res["nfev"], cc.last_static["nfev"])
    q = res["solution"]
    finite = np.all(np.isfinite(q), axis=1)

    # ---- 1a. collision + cost kernels on the GPU's own curves: 100 %
    full = ms.curves_batch(idx, theta, q)                                 # 4x4 curves of the same unknowns
    o_obs, o_goal, o_cost = orc.curves_eval_batch(oenv, full["g"], full["offsets"], 3, RAD, GOAL_WEIGHT)
    ok = finite & np.isfinite(o_obs)
    bad, n_band = sp.collision_contract(obs[ok], goal[ok], cost[ok], o_obs[ok], o_goal[ok], o_cost[ok])
    assert not bad.any(), (int(bad.sum()), np.flatnonzero(bad)[:10])
    assert np.array_equal(coll.astype(bool), obs < 0)
    # ---- 1b. curves kernel against the oracle's integrate_g of the same unknowns.  integrate_g evaluates the section
    # polynomials at ABSOLUTE node coordinates up to x_max = 99 (static.py:192-195), so a coefficient c of x^2 enters an
    # angle as c x^3 / 3: rounding alone moves the curve by ~ eps |c| x_max^3.  Ordinary unknowns (|c| < 1) stay at 1e-10;
    # the +-1e5 cancelling torsion coefficients that hybrd leaves in one-step sections are held to that bound
    sub = np.flatnonzero(finite)[:: max(1, B // 4000)]
    xmax = float(N[-1] - 1) * bench.DX
    T = 3
    worst_ratio = 0.0
    for b in sub:
        ref = np.concatenate(orc.static_curves(s, idx[b], theta[b], q[b]), 0)
        lo_, hi_ = full["offsets"][b * T], full["offsets"][(b + 1) * T]
        from util import rel_err
        tol = 1e-10 + 32.0 * np.finfo(float).eps * np.abs(q[b]).max() * xmax ** 3
        worst_ratio = max(worst_ratio, rel_err(full["g"][lo_:hi_], ref) / tol)
    assert worst_ratio <= 1.0, worst_ratio

    # ---- 2./3. solver against the oracle under the same one-step rule, and the oracle against itself
    o1 = orc.static_batch(s, idx, theta, one_step_rule=1, want_residual=True)
    rng = np.random.default_rng(7)
    x0 = orc.static_init_guess(s)
    x0p = x0[None, :] * (1.0 + 1e-13 * rng.standard_normal((B, x0.size)))
    o1p = orc.static_batch(s, idx, theta, one_step_rule=1, x0=x0p)
    bg = sp.buckets(q, res["info"], o1["solution"], o1["info"])
    bo = sp.buckets(o1p["solution"], o1p["info"], o1["solution"], o1["info"])
    report = {"workload": [lo, hi], "B": B, "one_step_fraction": float(short.mean()), "in_band": n_band,
              "curves_worst_error_over_bound": worst_ratio, "curves_checked": int(sub.size)}
    for name, sel in (("well_posed", ~short), ("one_step", short)):
        if not sel.any():
            continue
        rg, ro = sp.rates(bg, sel), sp.rates(bo, sel)
        report[name] = {"gpu_vs_oracle": rg, "oracle_vs_itself_1e-13": ro}
        # not in bucket A: bounded by the oracle's own sensitivity (twice its rate + 0.5 %)
        assert 1.0 - rg["A"] - rg["N"] <= 2.0 * (1.0 - ro["A"] - ro["N"]) + 0.005, (name, rg, ro)
        assert rg["F"] <= 2.0 * ro["F"] + 0.005, (name, rg, ro)
    # bucket B (and F): the GPU's unknowns are a root for the oracle's residual wherever the GPU reports convergence
    conv = (res["info"] == 1) & finite
    f_gpu = sp.residual_norms(orc, s, idx[conv], theta[conv], q[conv])
    f_orc = np.linalg.norm(o1["residual"], axis=1)[o1["info"] == 1]
    report["residual_norm"] = {"gpu_median": float(np.median(f_gpu)), "gpu_p99": float(np.quantile(f_gpu, 0.99)),
                               "gpu_max": float(f_gpu.max()), "oracle_median": float(np.median(f_orc)),
                               "oracle_p99": float(np.quantile(f_orc, 0.99)), "oracle_max": float(f_orc.max())}
    assert np.quantile(f_gpu, 0.99) <= 10.0 * np.quantile(f_orc, 0.99)
    assert np.median(f_gpu) <= 2.0 * np.median(f_orc)
    notA = conv.copy()
    notA[conv] = ~bg["A"][conv]
    if notA.any():
        f_b = sp.residual_norms(orc, s, idx[notA], theta[notA], q[notA])
        report["residual_norm"]["gpu_other_root_p99"] = float(np.quantile(f_b, 0.99))
        assert np.quantile(f_b, 0.99) <= 10.0 * np.quantile(f_orc, 0.99)
    # evaluation counts on the same path
    a = bg["A"]
    report["nfev"] = {"gpu_mean": float(res["nfev"].mean()), "oracle_mean": float(o1["nfev"].mean()),
                      "median_gap_same_root": float(np.median(np.abs(res["nfev"][a] - o1["nfev"][a])))}
    assert np.median(np.abs(res["nfev"][a] - o1["nfev"][a])) <= 1
    report["converged"] = {"gpu": float((res["info"] == 1).mean()), "oracle": float((o1["info"] == 1).mean())}

    # ---- 4. one-step sections against the reference's literal hybrd
    if short.any():
        o0 = orc.static_batch(s, idx[short], theta[short], one_step_rule=0, want_residual=True)
        o0p = orc.static_batch(s, idx[short], theta[short], one_step_rule=0, x0=x0p[short])
        b0 = sp.buckets(q[short], res["info"][short], o0["solution"], o0["info"])
        b00 = sp.buckets(o0p["solution"], o0p["info"], o0["solution"], o0["info"])
        all_ = np.ones(int(short.sum()), bool)
        report["one_step_vs_literal_hybrd"] = {"gpu_vs_literal": sp.rates(b0, all_), "literal_vs_itself_1e-13": sp.rates(b00, all_),
                                               "literal_converged": float((o0["info"] == 1).mean()),
                                               "literal_residual_p99": float(np.quantile(
                                                   np.linalg.norm(o0["residual"], axis=1)[o0["info"] == 1], 0.99))}
        # the converged flag is what survives of the comparison there
        assert sp.rates(b0, all_)["F"] <= 2.0 * sp.rates(b00, all_)["F"] + 0.02
    print("HEADLINE_PARITY " + json.dumps(report))
    out = ROOT / "gpurun_out"
    if out.is_dir():
        (out / f"headline_parity_{lo}_{hi}.json").write_text(json.dumps(report, indent=1))


def test_one_step_option_literal_qr_kernel():
    """CTRB_OPT_STATIC_ONE_STEP = 1: one-step sections go through MINPACK's own QR machinery (static_solve_kernel), the
    reference's literal treatment.  Against the oracle's literal hybrd: the converged flag agrees as often as the
    oracle agrees with itself under a 1e-13 perturbation, and the unknowns are roots of the oracle's residual."""
    bench, orc, ms, cc, s, oenv = _objects()
    B = 6000
    idx, theta = bench.make_inputs(B, 1003, 1, 8)
    short = sp.one_step(idx, np.asarray(bench.NPTS))
    idx, theta = idx[short], theta[short]
    n = idx.shape[0]
    with ms.ctx.options(static_one_step=1):
        res = ms.solve_g_batch(idx, theta, want_g=False)
        assert ms.ctx.static_fallbacks() == n                      # all of them through the QR kernel
    res_default = ms.solve_g_batch(idx, theta, want_g=False)
    assert ms.ctx.static_fallbacks() <= 0.02 * n                   # default: solved by the explicit-inverse kernel
    o0 = orc.static_batch(s, idx, theta, one_step_rule=0, want_residual=True)
    rng = np.random.default_rng(11)
    x0 = orc.static_init_guess(s)
    o0p = orc.static_batch(s, idx, theta, one_step_rule=0, x0=x0[None, :] * (1.0 + 1e-13 * rng.standard_normal((n, x0.size))))
    all_ = np.ones(n, bool)
    rg = sp.rates(sp.buckets(res["solution"], res["info"], o0["solution"], o0["info"]), all_)
    ro = sp.rates(sp.buckets(o0p["solution"], o0p["info"], o0["solution"], o0["info"]), all_)
    print("ONE_STEP_QR", json.dumps({"gpu_qr_vs_literal": rg, "literal_vs_itself": ro}))
    assert rg["F"] <= 2.0 * ro["F"] + 0.02, (rg, ro)
    conv = res["info"] == 1
    f_gpu = sp.residual_norms(orc, s, idx[conv], theta[conv], res["solution"][conv])
    f_orc = np.linalg.norm(o0["residual"], axis=1)[o0["info"] == 1]
    assert np.quantile(f_gpu, 0.99) <= 10.0 * np.quantile(f_orc, 0.99)
    assert np.all(res["info"] >= 1) and np.all(res_default["info"] >= 1)


def test_one_step_rule_in_both_solve_kernels():
    """The reduced-system rule for one-step sections is carried by both solve kernels: the explicit-inverse kernel
    (default) and MINPACK's QR machinery (ctx option static_solver = 1, which also receives the few configurations whose
    Jacobian turns out singular at a re-evaluation).  Same rule, two factorizations: against the oracle's hybrd on the
    reduced system each lands on the same root as often as the oracle does under a 1e-13 perturbation of itself."""
    bench, orc, ms, cc, s, oenv = _objects()
    idx, theta = bench.make_inputs(12_000, 1003, 1, 8)
    short = sp.one_step(idx, np.asarray(bench.NPTS))
    idx, theta = idx[short], theta[short]
    n = idx.shape[0]
    fast = ms.solve_g_batch(idx, theta, want_g=False)
    reasons = ms.ctx.static_fallback_reasons()
    assert sum(reasons.values()) == ms.ctx.static_fallbacks() and reasons["degenerate_section"] == 0
    with ms.ctx.options(static_solver=1):
        qr = ms.solve_g_batch(idx, theta, want_g=False)
    o1 = orc.static_batch(s, idx, theta, one_step_rule=1)
    x0 = orc.static_init_guess(s)
    o1p = orc.static_batch(s, idx, theta, one_step_rule=1,
                           x0=x0[None, :] * (1.0 + 1e-13 * np.random.default_rng(3).standard_normal((n, x0.size))))
    all_ = np.ones(n, bool)
    ro = sp.rates(sp.buckets(o1p["solution"], o1p["info"], o1["solution"], o1["info"]), all_)
    for name, res in (("fast", fast), ("qr", qr)):
        rg = sp.rates(sp.buckets(res["solution"], res["info"], o1["solution"], o1["info"]), all_)
        print("ONE_STEP_RULE", name, json.dumps(rg), "oracle vs itself", json.dumps(ro))
        assert 1.0 - rg["A"] - rg["N"] <= 2.0 * (1.0 - ro["A"] - ro["N"]) + 0.005, (name, rg, ro)
        # the pinned x^2 coefficients keep their initial value exactly
        pins = [2, 5, 8, 17, 20, 23]
        e = np.asarray(bench.NPTS)[None, :] - 1 - idx
        for sec, cols in ((0, pins[:3]), (1, pins[3:])):
            sel = e[:, sec] == 1
            assert np.all(res["solution"][sel][:, cols] == x0[cols])


def test_position_curve_kernels_agree():
    """ctrb_static_eval_batch integrates node positions with one thread per configuration (static_curves_pos_kernel);
    the warp-per-configuration scan kernel (ctx option static_curves_scan) multiplies the same Magnus steps in another
    order.  Same collision bits, distances to 1e-9 mm, on both exposure distributions."""
    bench, orc, ms, cc, s, oenv = _objects()
    for lo, hi, B in ((1, 8, 80_000), (1, 33, 70_000)):      # >= 65536: below that the library takes the scan kernel itself
        idx, theta = bench.make_inputs(B, 2024, lo, hi)
        a = cc.evaluate_batch(ms, idx, theta, RAD, goal_weight=GOAL_WEIGHT)
        with cc.ctx.options(static_curves_scan=1):
            b = cc.evaluate_batch(ms, idx, theta, RAD, goal_weight=GOAL_WEIGHT)
        band = (np.abs(a[0]) < 1e-9) | (np.abs(b[0]) < 1e-9)
        assert np.array_equal((a[0] < 0)[~band], (b[0] < 0)[~band])
        free = (a[0] > 0) & (b[0] > 0)
        # huge cancelling torsion coefficients (one-step sections) amplify the rounding of either order: see 1b above
        big = np.abs(ms.solve_g_batch(idx, theta, want_g=False)["solution"]).max(1) > 1e2
        assert np.max(np.abs(a[0][free & ~big] - b[0][free & ~big])) < 1e-9
        assert np.max(np.abs(a[1][~big] - b[1][~big])) < 1e-9
        assert np.max(np.abs(a[0][free] - b[0][free])) < 1e-3


def test_integrate_g_from_given_unknowns():
    """ctrb_static_curves_batch: integrate_g (static.py:165-198, with its absolute-index quirk) from GIVEN unknowns,
    against the reference's own curves of its own solutions (tests/golden/static.npz) at the contract's 1e-10, and
    against the oracle's integrate_g on random unknowns -- one-step sections included: given the same q there is
    nothing ill-posed about the integration."""
    from conftest import GOLDEN
    from models_def import STATIC_MODELS
    from util import gpu_static, orc_static, rel_err
    from oracle import oracle as orc
    gs = np.load(GOLDEN / "static.npz")
    for name in STATIC_MODELS:
        m, s = gpu_static(name), orc_static(name)
        T = m.tube_num
        cases = STATIC_MODELS[name][-1]
        for c, (idx, th) in enumerate(cases):
            # the reference's own solutions of one-step sections carry +-3.5e3 cancelling coefficients evaluated at
            # x ~ 119: rounding is amplified ~1e10 there, and the oracle itself reproduces the
            # reference's curves to 1e-4 only (tests/test_oracle_static.py); everywhere else the contract's 1e-10 holds
            short = min(m.num_discrete_points[n] - idx[n] for n in range(T)) <= 2
            for tag in ("tight", "default"):
                ref_q, ref_g = gs[f"{name}/c{c}/sol_{tag}"], gs[f"{name}/c{c}/g_{tag}"]
                if not np.all(np.isfinite(ref_g)):
                    continue
                got = m.curves_batch([idx], [th], ref_q)
                assert got["g"].shape == ref_g.shape
                assert rel_err(got["g"], ref_g) <= (1e-4 if short else 1e-10), (name, c, tag, rel_err(got["g"], ref_g))
        # random unknowns near the initial guess, all exposures (one-step and zero-length sections included)
        rng = np.random.default_rng(5)
        N = np.asarray(m.num_discrete_points)
        B = 200
        idx = np.stack([rng.integers(0, N[t], B) for t in range(T)], 1).astype(np.int32)
        idx[:20] = N[None, :] - 2
        th = rng.uniform(0, 2 * np.pi, (B, T))
        q = m.general_init_guess[None, :] * (1.0 + 0.3 * rng.standard_normal((B, m.num_unknowns))) \
            + 1e-4 * rng.standard_normal((B, m.num_unknowns))
        got = m.curves_batch(idx, th, q)
        for b in range(B):
            ref = np.concatenate(orc.static_curves(s, idx[b], th[b], q[b]), 0)
            lo, hi = got["offsets"][b * T], got["offsets"][(b + 1) * T]
            assert rel_err(got["g"][lo:hi], ref) <= 1e-10, (name, b, rel_err(got["g"][lo:hi], ref))


def test_static_eval_device_mode_does_not_synchronise():
    """include/ctrb200.h: a CTRB_MEM_DEVICE call is stream-ordered and returns without synchronising.  A long kernel
    is queued on the ctx's stream first; ctrb_static_eval_batch must return while it is still running."""
    import ctypes as C
    import time
    import torch
    from ctrdapp_b200 import _lib as L
    bench, orc, ms, cc, s, oenv = _objects()
    dev = torch.device("cuda", 0)
    B = 20_000
    idx, theta = bench.make_inputs(B, 5, 2, 8)
    idx_d, th_d = torch.from_numpy(idx).to(dev), torch.from_numpy(theta).to(dev)
    outs = [torch.empty(B, dtype=torch.float64, device=dev) for _ in range(3)] + [torch.empty(B, dtype=torch.uint8, device=dev)] \
        + [torch.empty(B, dtype=torch.int32, device=dev) for _ in range(2)]
    cd = L.CostDesc(L.COST_TYPES["square_obstacle_avg_plus_weighted_goal"], GOAL_WEIGHT, 1, 0.0)
    rad = np.ascontiguousarray(RAD, np.float64)
    ctx = ms.ctx
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)
    try:
        def call():
            L.check(L.lib().ctrb_static_eval_batch(ctx.handle, ms.handle, cc.handle, C.byref(cd), B, L.ptr(idx_d), L.ptr(th_d),
                                                   L.ptr(rad), *[L.ptr(o) for o in outs], L.MEM_DEVICE))
        call()                                   # warm-up: scratch buffers grow here (growth may synchronise)
        torch.cuda.synchronize()
        ref = [o.clone() for o in outs]
        with torch.cuda.stream(stream):
            torch.cuda._sleep(int(2.0e9))        # ~1 s of device time queued in front of the call
            t0 = time.perf_counter()
            call()
            host_s = time.perf_counter() - t0
            busy = not stream.query()
        torch.cuda.synchronize()
        assert busy and host_s < 0.25, (busy, host_s)
        for a, b in zip(ref, outs):
            assert torch.equal(a, b) or torch.allclose(a, b, equal_nan=True)
    finally:
        ctx.set_stream(None)


def test_static_host_memory_pipeline_equals_device_path():
    """ctrb_static_eval_batch(CTRB_MEM_HOST) on 2^22 or more configurations runs as a copy / compute pipeline over chunks
    (two streams, two sets of staging buffers): identical outputs to the single-sequence device path; a bad index in a late
    chunk is still a ValueError that leaves the context usable."""
    import ctypes as C
    import torch
    from ctrdapp_b200 import _lib as L
    bench, orc, ms, cc, s, oenv = _objects()
    dev = torch.device("cuda", 0)
    B = (1 << 22) + 100_000                    # three chunks of 2^21, 2^21 and 100 000 (>= 65 536: the same curves kernel)
    idx, theta = bench.make_inputs(B, 77, 1, 8)
    a = cc.evaluate_batch(ms, idx, theta, RAD, goal_weight=GOAL_WEIGHT)          # host buffers: pipelined
    info_a, nfev_a = cc.last_static["info"].copy(), cc.last_static["nfev"].copy()
    idx_d, th_d = torch.from_numpy(idx).to(dev), torch.from_numpy(theta).to(dev)
    outs = [torch.empty(B, dtype=torch.float64, device=dev) for _ in range(3)] + [torch.empty(B, dtype=torch.uint8, device=dev)] \
        + [torch.empty(B, dtype=torch.int32, device=dev) for _ in range(2)]
    cd = L.CostDesc(L.COST_TYPES["square_obstacle_avg_plus_weighted_goal"], GOAL_WEIGHT, 1, 0.0)
    rad = np.ascontiguousarray(RAD, np.float64)
    L.check(L.lib().ctrb_static_eval_batch(ms.ctx.handle, ms.handle, cc.handle, C.byref(cd), B, L.ptr(idx_d), L.ptr(th_d),
                                           L.ptr(rad), *[L.ptr(o) for o in outs], L.MEM_DEVICE))
    ms.ctx.synchronize()
    for host, d in zip(list(a) + [info_a, nfev_a], outs):
        assert np.array_equal(host, d.cpu().numpy(), equal_nan=True)
    bad = idx.copy()
    bad[B - 5, 1] = 10_000
    with pytest.raises(ValueError, match="outside"):
        cc.evaluate_batch(ms, bad, theta, RAD, goal_weight=GOAL_WEIGHT)
    small = cc.evaluate_batch(ms, idx[:1000], theta[:1000], RAD, goal_weight=GOAL_WEIGHT)   # the context still works
    assert np.array_equal(small[3], a[3][:1000])         # (a batch this small takes the scan curves kernel: last-bit differences)
    np.testing.assert_allclose(small[0], a[0][:1000], rtol=1e-9)
