"""GPU parity: the static model kernels (closed-form residual, warp-cooperative MINPACK hybrd,
scan-integrated curves) vs the reference's golden vectors and vs the literal CPU oracle."""
import numpy as np
import pytest

from conftest import GOLDEN, FIXTURES
from models_def import STATIC_MODELS
from util import gpu_static, orc_static, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gs():
    return np.load(GOLDEN / "static.npz")


def _short(m, idx):
    return min(m.num_discrete_points[n] - idx[n] for n in range(m.tube_num)) <= 2


@pytest.mark.parametrize("name", list(STATIC_MODELS))
def test_residual_golden(gs, name):
    """static_equilibrium(x): 4 bracketing nodes + closed forms vs the reference's per-node march."""
    m = gpu_static(name)
    np.testing.assert_array_equal(m.general_init_guess, gs[f"{name}/x0"])
    cases = STATIC_MODELS[name][-1]
    for c, (idx, th) in enumerate(cases):
        xs, Fs = gs[f"{name}/c{c}/res_x"], gs[f"{name}/c{c}/res_F"]
        got = m.static_equilibrium_batch([idx] * len(xs), [th] * len(xs), xs)
        for F, G in zip(Fs, got):
            assert np.max(np.abs(G - F)) <= 1e-10 * max(np.max(np.abs(F)), 1e-300), (name, c)


@pytest.mark.parametrize("name", list(STATIC_MODELS))
def test_curves_from_reference_solution(gs, name):
    """integrate_g parity given the SAME q: feed the reference's solution as the initial guess with a
    loose xtol so that hybrd accepts it after its first Jacobian (q stays within ~1e-9 of the input)."""
    m = gpu_static(name)
    cases = STATIC_MODELS[name][-1]
    for c, (idx, th) in enumerate(cases):
        if gs[f"{name}/c{c}/meta_tight"][1] == 0:
            continue
        ref_q = gs[f"{name}/c{c}/sol_tight"]
        res = m.solve_g_batch([idx], [th], initial_guess=ref_q, xtol=1e-13)
        ref_g = gs[f"{name}/c{c}/g_tight"]
        assert res["g"].shape == ref_g.shape
        if _short(m, idx):
            continue  # degenerate one-step sections: compared through the oracle only
        assert np.max(np.abs(res["solution"][0] - ref_q)) <= 1e-9 * max(1.0, np.max(np.abs(ref_q)))
        assert rel_err(res["g"], ref_g) < 1e-8, (name, c, rel_err(res["g"], ref_g))


@pytest.mark.parametrize("name", list(STATIC_MODELS))
def test_solve_g_golden(gs, name):
    """Static.solve_g from the default initial guess: same root as SciPy's hybr, same success flag, a
    comparable evaluation count; g-curves within the reference's own loose-vs-tight deviation."""
    m = gpu_static(name)
    cases = STATIC_MODELS[name][-1]
    idx = np.array([c[0] for c in cases], np.int32)
    th = np.array([c[1] for c in cases])
    res = m.solve_g_batch(idx, th)
    T = m.tube_num
    for c in range(len(cases)):
        nfev_ref, ok_ref, _ = gs[f"{name}/c{c}/meta_default"]
        assert bool(res["converged"][c]) == bool(ok_ref), (name, c, res["info"][c])
        if not ok_ref or _short(m, cases[c][0]):
            continue
        ref_q = gs[f"{name}/c{c}/sol_tight"]
        assert np.max(np.abs(res["solution"][c] - ref_q)) <= 5e-6 * max(1.0, np.max(np.abs(ref_q))), (name, c)
        assert abs(int(res["nfev"][c]) + 2 - nfev_ref) <= 0.4 * nfev_ref, (name, c, res["nfev"][c], nfev_ref)
        lo, hi = res["offsets"][c * T], res["offsets"][(c + 1) * T]
        ref_g = gs[f"{name}/c{c}/g_default"]
        dev_ref = rel_err(gs[f"{name}/c{c}/g_tight"], ref_g)          # the reference's own loose-tolerance deviation
        # both stop when the step falls below xtol = 1.49e-8 (relative): curves agree to the solver tolerance
        assert rel_err(res["g"][lo:hi], ref_g) <= max(1e-6, 20 * dev_ref), (name, c)
    # reference-shaped call
    g = m.solve_g(list(cases[0][0]), list(cases[0][1]))
    assert [len(t) for t in g] == gs[f"{name}/c0/counts_default"].tolist()
    assert m.last_solution["success"]


def _bucket_check(m, s, idx, th, one_step_rule, seed, label):
    """GPU solve vs the oracle's hybrd, per bucket (tests/static_parity.py), bounded by the oracle's OWN sensitivity: the
    same comparison of the oracle with itself from an initial guess perturbed by 1e-13, measured here.  No fixed
    threshold below 1: where the oracle is insensitive, every configuration has to land on its root."""
    import static_parity as sp
    from oracle import oracle as orc
    B = idx.shape[0]
    res = m.solve_g_batch(idx, th)
    o = orc.static_batch(s, idx, th, one_step_rule=one_step_rule, want_residual=True)
    x0 = orc.static_init_guess(s)
    x0p = x0[None, :] * (1.0 + 1e-13 * np.random.default_rng(seed).standard_normal((B, x0.size)))
    op = orc.static_batch(s, idx, th, one_step_rule=one_step_rule, x0=x0p)
    allc = np.ones(B, bool)
    rg = sp.rates(sp.buckets(res["solution"], res["info"], o["solution"], o["info"]), allc)
    ro = sp.rates(sp.buckets(op["solution"], op["info"], o["solution"], o["info"]), allc)
    print(f"{label}: gpu vs oracle {rg}; oracle vs itself (1e-13) {ro}")
    assert 1.0 - rg["A"] - rg["N"] <= 2.0 * (1.0 - ro["A"] - ro["N"]) + 0.005, (label, rg, ro)
    assert rg["F"] <= 2.0 * ro["F"] + 0.005, (label, rg, ro)
    # whatever the bucket: a converged GPU solve is a root of the oracle's marched residual
    conv = res["converged"] & np.all(np.isfinite(res["solution"]), axis=1)
    f_gpu = sp.residual_norms(orc, s, idx[conv], th[conv], res["solution"][conv])
    f_orc = np.linalg.norm(o["residual"], axis=1)[o["info"] == 1]
    assert np.quantile(f_gpu, 0.99) <= 10.0 * np.quantile(f_orc, 0.99) and np.median(f_gpu) <= 2.0 * np.median(f_orc)
    # same root => same curves to the solver tolerance, same evaluation count
    same = sp.buckets(res["solution"], res["info"], o["solution"], o["info"])["A"]
    assert np.median(np.abs(res["nfev"][same] - o["nfev"][same])) <= 1
    return res, o, same


def test_random_batch_vs_oracle():
    """300 random C1 configurations (two tubes, configuration/config.yaml as a static model), >= 3 nodes per section."""
    from oracle import oracle as orc
    name = "c1_two_tube"
    m, s = gpu_static(name), orc_static(name)
    rng = np.random.default_rng(1001)
    B = 300
    idx = rng.integers(0, 118, (B, 2)).astype(np.int32)   # >= 3 nodes per section
    th = rng.uniform(0, 2 * np.pi, (B, 2))
    res, o, same = _bucket_check(m, s, idx, th, 0, 1, "C1 well-posed")
    worst_g = 0.0
    for b in np.flatnonzero(same)[::3]:
        g = orc.static_curves(s, idx[b], th[b], o["solution"][b])
        lo, hi = res["offsets"][b * 2], res["offsets"][b * 2 + 2]
        worst_g = max(worst_g, rel_err(res["g"][lo:hi], np.concatenate(g, 0)))
    assert worst_g < 1e-4, worst_g  # both stopped at xtol = 1.49e-8: curves agree to the solver tolerance


def test_random_three_tube_batch_vs_oracle():
    """600 random C3 configurations (BASELINE configs[2]'s model), exposure 2-12 nodes per tube, then the same batch with
    one-step sections mixed in: by default those are solved on the reduced system (against the oracle's hybrd on that
    system), with ctx option static_one_step = 1 they go to the QR kernel, MINPACK's literal treatment (against the
    oracle's literal hybrd, where only the converged flag can be compared: its null-space component is rounding noise)."""
    from oracle import oracle as orc
    name = "c3_three_tube"
    m, s = gpu_static(name), orc_static(name)
    N = np.asarray(m.num_discrete_points)
    rng = np.random.default_rng(2003)
    B = 600
    e = rng.integers(2, 13, (B, 3))
    idx = (N[None, :] - 1 - e).astype(np.int32)
    th = rng.uniform(0, 2 * np.pi, (B, 3))
    _bucket_check(m, s, idx, th, 0, 2, "C3 well-posed")
    assert m.ctx.static_fallbacks() <= 0.02 * B          # a numerically singular Jacobian is the exception here
    e[::3, rng.integers(0, 2)] = 1
    idx = (N[None, :] - 1 - e).astype(np.int32)
    _bucket_check(m, s, idx, th, 1, 3, "C3 with one-step sections, reduced-system rule")
    assert m.ctx.static_fallbacks() <= 0.02 * B
    with m.ctx.options(static_one_step=1):
        res = m.solve_g_batch(idx, th)
        assert m.ctx.static_fallbacks() >= B // 3
    assert np.all(res["info"] >= 1) and np.all(res["nfev"] >= 31) and np.all(np.isfinite(res["solution"]))
    sub = np.arange(0, B, 3)
    o0 = orc.static_batch(s, idx[sub], th[sub], one_step_rule=0)
    x0 = orc.static_init_guess(s)
    o0p = orc.static_batch(s, idx[sub], th[sub], one_step_rule=0,
                           x0=x0[None, :] * (1.0 + 1e-13 * np.random.default_rng(4).standard_normal((sub.size, x0.size))))
    flag_gpu = np.mean((o0["info"] == 1) != res["converged"][sub])
    flag_self = np.mean((o0["info"] == 1) != (o0p["info"] == 1))
    assert flag_gpu <= 2.0 * flag_self + 0.02, (flag_gpu, flag_self)


def test_static_fused_eval():
    """BASELINE configs[0]: 2-tube static model, 1024 random configs vs in_simple.STL (C1, all collide at the
    base) and vs the translated box (C1b, distances exercised): static solve -> collision -> cost."""
    from oracle import oracle as orc
    from ctrdapp_b200.collision import CollisionChecker
    name = "c1_two_tube"
    m, s = gpu_static(name), orc_static(name)
    rng = np.random.default_rng(1001)
    B = 1024
    idx = rng.integers(0, 120, (B, 2)).astype(np.int32)
    th = rng.uniform(0, 2 * np.pi, (B, 2))
    cc = CollisionChecker(FIXTURES / "env_c1_in_simple.json")
    obs, goal, cost, coll = cc.evaluate_batch(m, idx, th, [0.9, 0.6], goal_weight=3.0)
    assert np.all(obs == -1.0) and np.all(np.isnan(cost)) and cc.last_static["converged"].mean() > 0.9
    cc = CollisionChecker(FIXTURES / "env_boxes_multi.json")
    oenv = orc.Env.from_json(FIXTURES / "env_boxes_multi.json")
    obs, goal, cost, coll = cc.evaluate_batch(m, idx, th, [0.9, 0.6], goal_weight=3.0)
    # the solve is deterministic: the fused path's curves are those of solve_g_batch, so the oracle's collision
    # query on those curves must agree to the collision contract (1e-6 mm), whatever equilibrium was reached
    res = m.solve_g_batch(idx, th)
    assert np.array_equal(res["info"], cc.last_static["info"])
    checked = 0
    for b in range(0, B, 8):
        curve = [res["g"][res["offsets"][b * 2 + n]:res["offsets"][b * 2 + n + 1]] for n in range(2)]
        if not np.all(np.isfinite(res["g"][res["offsets"][b * 2]:res["offsets"][b * 2 + 2]])):
            continue
        o_obs, o_goal = oenv.check_collision(curve, [0.9, 0.6])
        assert (o_obs < 0) == (obs[b] < 0)
        if o_obs > 0:
            assert abs(o_obs - obs[b]) < 1e-6 and abs(o_goal - goal[b]) < 1e-6
            assert cost[b] == pytest.approx(3.0 * max(o_goal, 0.0) + 1.0 / o_obs ** 2, rel=1e-9)
        checked += 1
    assert checked > 100



@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c1_two_tube", "c3_three_tube"])
def test_fast_kernel_instantiations_agree(name):
    """static_fast_kernel is instantiated per system size (27 / 30 unknowns for three tubes, 12 / 15 for two; loops
    unrolled) with a run-time-n instantiation for every other size: both must give the same iterates bit for bit, from
    the default initial guess (reduced system) and from a caller-supplied one (full system)."""
    m = gpu_static(name)
    N = np.asarray(m.num_discrete_points)
    T = m.tube_num
    rng = np.random.default_rng(3003)
    B = 400
    e = np.stack([rng.integers(2, min(12, N[t] - 1) + 1, B) for t in range(T)], 1)
    idx = (N[None, :] - 1 - e).astype(np.int32)
    th = rng.uniform(0, 2 * np.pi, (B, T))
    guess = np.tile(m.general_init_guess, (B, 1)) * (1.0 + 1e-3 * rng.standard_normal((B, m.num_unknowns)))
    for x0 in (None, guess):
        a = m.solve_g_batch(idx, th, initial_guess=x0)
        fb_a = m.ctx.static_fallbacks()
        with m.ctx.options(fast_generic_n=1):
            b = m.solve_g_batch(idx, th, initial_guess=x0)
            assert m.ctx.static_fallbacks() == fb_a
        assert fb_a <= 0.05 * B  # the batch is solved by the fast kernel, not handed to the QR kernel
        np.testing.assert_array_equal(a["info"], b["info"])
        np.testing.assert_array_equal(a["nfev"], b["nfev"])
        np.testing.assert_array_equal(a["solution"], b["solution"])
        np.testing.assert_array_equal(a["g"], b["g"])
        assert a["converged"].mean() > 0.9


@pytest.mark.gpu
def test_block_structured_inverse_vs_gauss_jordan():
    """Three tubes, default initial guess: static_fast_kernel inverts the block-tridiagonal forward-difference Jacobian
    through its 12 | 3 | 12 structure (fast_invert_block27).  Against the general Gauss-Jordan inverse of the same
    kernel (ctx option fast_block_invert = 0) the inverse differs by rounding only: same hand-over set, and the same flag,
    evaluation count and root in all but the few configurations where hybrd's path is sensitive to the last bit
    (the CPU port itself changes root in ~4 % of cases under a 1e-13 perturbation)."""
    m = gpu_static("c3_three_tube")
    N = np.asarray(m.num_discrete_points)
    rng = np.random.default_rng(4004)
    B = 2000
    e = rng.integers(2, 13, (B, 3))
    idx = (N[None, :] - 1 - e).astype(np.int32)
    th = rng.uniform(0, 2 * np.pi, (B, 3))
    a = m.solve_g_batch(idx, th, want_g=False)
    fb_a = m.ctx.static_fallbacks()
    with m.ctx.options(fast_block_invert=0):
        b = m.solve_g_batch(idx, th, want_g=False)
        assert m.ctx.static_fallbacks() == fb_a
    same_flag = (a["info"] == b["info"]).mean()
    same_nfev = (a["nfev"] == b["nfev"]).mean()
    both = a["converged"] & b["converged"]
    scale = np.maximum(1.0, np.abs(b["solution"]).max(axis=1))
    close = (np.abs(a["solution"] - b["solution"]).max(axis=1) <= 1e-6 * scale)[both].mean()
    # measured: same info 1.0000, same nfev 0.9975, same root 1.0000
    assert same_flag >= 0.98 and same_nfev >= 0.95 and close >= 0.98, (same_flag, same_nfev, close)
    print("block vs Gauss-Jordan: same info %.4f, same nfev %.4f, same root %.4f" % (same_flag, same_nfev, close))
    assert a["converged"].mean() > 0.95 and abs(a["converged"].mean() - b["converged"].mean()) < 0.01


def test_static_empty_and_invalid_batches():
    """B = 0 through every static entry point (the reference has no batch; an empty list of candidates is what a
    planner round without survivors hands over), and indices outside the tubes raise like the reference's IndexError path."""
    from ctrdapp_b200.collision import CollisionChecker
    m = gpu_static("c3_three_tube")
    nq = m.num_unknowns
    idx0, th0 = np.zeros((0, 3), np.int32), np.zeros((0, 3))
    res = m.solve_g_batch(idx0, th0)
    assert res["g"].shape == (0, 4, 4) and res["offsets"].tolist() == [0]
    assert res["solution"].shape == (0, nq) and res["nfev"].shape == (0,) and res["info"].shape == (0,)
    assert m.curves_batch(idx0, th0, np.zeros((0, nq)))["g"].shape == (0, 4, 4)
    assert m.static_equilibrium_batch(idx0, th0, np.zeros((0, nq))).shape == (0, nq)
    cc = CollisionChecker(FIXTURES / "env_c3_colon_straight.json", ctx=m.ctx)
    out = cc.evaluate_batch(m, idx0, th0, [1.2, 0.9, 0.6])
    assert all(len(o) == 0 for o in out) and cc.last_static["info"].shape == (0,)
    with pytest.raises(ValueError):
        m.solve_g_batch([[0, 0, 500]], [[0.0, 0.0, 0.0]])
    with pytest.raises(ValueError):
        cc.evaluate_batch(m, [[-1, 0, 0]], [[0.0, 0.0, 0.0]], [1.2, 0.9, 0.6])
