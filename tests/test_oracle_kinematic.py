"""CPU: the C oracle restatement vs golden vectors produced by the Python reference
(tests/golden/make_golden.py) and vs the reference's own unit-test literals."""
import math

import numpy as np
import pytest

from oracle import oracle as orc
from models_def import MODELS


def _split(flat, counts):
    out, o = [], 0
    for c in counts:
        out.append(flat[o:o + c])
        o += c
    return out


@pytest.mark.parametrize("name", list(MODELS))
def test_solve_g_matches_reference(golden_kin, name):
    bases, dof, q, lengths, dx = MODELS[name]
    m = orc.make_model(bases, dof, q, lengths, dx)
    assert [m.npts[n] for n in range(m.tube_num)] == golden_kin[f"{name}/npts"].tolist()
    flat = golden_kin[f"{name}/g_flat"]
    o = 0
    for idx, th, full, cnt in zip(golden_kin[f"{name}/g_idx"], golden_kin[f"{name}/g_theta"],
                                  golden_kin[f"{name}/g_full"], golden_kin[f"{name}/g_counts"]):
        g = orc.solve_g(m, idx.tolist(), th.tolist(), full=bool(full))
        assert [len(t) for t in g] == cnt.tolist()
        got = np.concatenate(g, 0)
        ref = flat[o:o + cnt.sum()]
        o += cnt.sum()
        scale = np.maximum(1.0, np.linalg.norm(ref[:, 0:3, 3], axis=1))[:, None, None]
        assert np.max(np.abs(got - ref) / scale) < 1e-12


@pytest.mark.parametrize("name", list(MODELS))
def test_solve_integrate_matches_reference(golden_kin, name):
    bases, dof, q, lengths, dx = MODELS[name]
    m = orc.make_model(bases, dof, q, lengths, dx)
    for c in range(4):
        k = f"{name}/si{c}"
        prev_g = _split(golden_kin[k + "/prev_g"], golden_kin[k + "/prev_counts"])
        d_th, d_ins, th, ins = golden_kin[k + "/args"]
        g, eta, nidx, tins, ftl = orc.solve_integrate(m, d_th, d_ins, th, ins, prev_g)
        assert nidx == golden_kin[k + "/new_idx"].tolist()
        np.testing.assert_allclose(tins, golden_kin[k + "/true_ins"], rtol=0, atol=1e-12)
        assert [len(t) for t in g] == golden_kin[k + "/g_counts"].tolist()
        np.testing.assert_allclose(np.concatenate(g, 0), golden_kin[k + "/g"], rtol=0, atol=2e-11)
        np.testing.assert_allclose(np.array([e[0] for e in eta]), golden_kin[k + "/eta"], rtol=1e-11, atol=1e-11)
        assert [len(t) for t in ftl] == golden_kin[k + "/ftl_counts"].tolist()
        np.testing.assert_allclose(np.concatenate(ftl, 0), golden_kin[k + "/ftl"], rtol=1e-10, atol=1e-10)
        f = np.concatenate(ftl, 0)
        mags = [orc.lib().orc_ftl_magnitude(orc._dp(np.ascontiguousarray(x)), 0) for x in f]
        tmags = [orc.lib().orc_ftl_magnitude(orc._dp(np.ascontiguousarray(x)), 1) for x in f]
        want = golden_kin[k + "/ftl_mag"]
        np.testing.assert_allclose([mags[-1], np.mean(mags), tmags[-1], np.mean(tmags)], want, rtol=1e-10)


def test_calculate_indices_golden(golden_kin):
    for row, want in zip(golden_kin["calc_idx/in"], golden_kin["calc_idx/out"]):
        p, n = orc.calculate_indices(row[0:2], row[2:4], row[4:6], float(row[6]))
        assert p + n == want.tolist(), row


def test_calculate_indices_reference_table():
    """test/test_model.py:353-393 literals."""
    def chk(prev, nxt, p_out, n_out, dx=1.0):
        p, n = orc.calculate_indices([prev], [nxt], [50], dx)
        assert (p, n) == ([p_out], [n_out]), (prev, nxt)
    for a in (-3, -1, 0):
        for b in (-3, -1, 0):
            chk(a, b, 0, 0)
    for a in (50, 51, 53):
        for b in (50, 51, 53):
            chk(a, b, 50, 50)
    for args in [(0, 0.49, 0, 1), (0.49, 0, 1, 0), (-5, 0.49, 0, 1), (0.49, -0.02, 1, 0), (49.51, 50, 49, 50),
                 (50, 49.51, 50, 49), (49.4, 50.3, 49, 50), (60, 49.05, 50, 49), (10.2, 10.3, 10, 11),
                 (10.3, 10.2, 11, 10), (19.9, 21.95, 20, 22), (21.95, 19.9, 22, 20), (19.49, 21.51, 19, 21),
                 (21.51, 19.49, 22, 20), (10, 10, 10, 10)]:
        chk(*args)
    chk(10.27, 10.31, 103, 104, dx=0.1)
    chk(10.31, 10.27, 103, 102, dx=0.1)


def test_exponential_map_golden(golden_kin):
    for row, want in zip(golden_kin["expmap/in"], golden_kin["expmap/out"]):
        got = orc.exponential_map(row[0], row[1:].reshape(4, 4))
        np.testing.assert_allclose(got.ravel(), want, rtol=1e-13, atol=1e-15)


def test_exponential_map_reference_literals():
    """test/test_model.py:77-81: theta = 0 exact, theta = pi to 3 decimals."""
    hat = np.array([[0, -3, 2, 4], [3, 0, -1, 5], [-2, 1, 0, 6], [0, 0, 0, 0.0]])
    np.testing.assert_equal(orc.exponential_map(0.0, hat),
                            np.array([[1, -3, 2, 4], [3, 1, -1, 5], [-2, 1, 1, 6], [0, 0, 0, 1.0]]))
    exp_pi = np.array([[-1.6344, 1.6608, -0.2291, 0.9604], [-0.8502, -1.0264, 1.6344, 5.6079],
                       [1.4449, 0.7974, -0.0132, 6.6079], [0, 0, 0, 1]])
    np.testing.assert_almost_equal(orc.exponential_map(math.pi, hat), exp_pi, decimal=3)


def test_quarter_circle_literals():
    """test/test_model.py:161-204."""
    m1 = orc.make_model(["linear"], [2], [1, 0], [math.pi / 2], math.pi / 20)
    g = orc.solve_g(m1, [0], None, full=True)[0]
    np.testing.assert_equal(g[0], np.eye(4))
    tip = np.array([[0, 0, 1, 1], [0, 1, 0, 0], [-1, 0, 0, -1], [0, 0, 0, 1.0]])
    np.testing.assert_almost_equal(g[-1], tip)
    m2 = orc.make_model(["linear", "linear"], [2, 2], [1, 0, 1, 0], [math.pi / 2] * 2, math.pi / 20)
    g = orc.solve_g(m2, [0, 0], [0, 0], full=True)
    np.testing.assert_almost_equal(g[1][-1], np.array([[-1, 0, 0, 0], [0, 1, 0, 0], [0, 0, -1, -2], [0, 0, 0, 1.0]]))
    g = orc.solve_g(m2, [0, 0], [0, math.pi / 2], full=True)
    np.testing.assert_almost_equal(g[1][-1], np.array([[0, 1, 0, 1], [1, 0, 0, 1], [0, 0, -1, -2], [0, 0, 0, 1.0]]))


def test_survey_known_answers():
    """SURVEY.md 8c probe values of the reference (17 digits)."""
    bases, dof, q, lengths, dx = MODELS["config_yaml"]
    m = orc.make_model(bases, dof, q, lengths, dx)
    g = orc.solve_g(m, [60, 30], [0.3, 1.2], full=False)
    assert [len(t) for t in g] == [61, 91]
    np.testing.assert_allclose(g[1][-1][0:3, 3], [7.422632086230158, 10.526089150858443, 11.508004938100957],
                               rtol=1e-12)
    np.testing.assert_allclose(g[1][-1][0, 0:3], [0.17737191758180187, -0.9819876527256568, -0.06511108006867973],
                               rtol=1e-12)
    g = orc.solve_g(m, [0, 0], [0.3, 1.2], full=True)
    np.testing.assert_allclose(g[1][-1][0:3, 3], [19.84874458988537, 40.760900602498936, 7.363386333942837], rtol=1e-12)
    prev = orc.solve_g(m, [62, 120], [0.25, 1.1], full=False)
    _, eta, nidx, tins, ftl = orc.solve_integrate(m, [.05, .1], [2, 0], [.3, 1.2], [60, 0], prev)
    assert nidx == [60, 120] and tins == [60, 0]
    np.testing.assert_allclose(eta[1][0], [0.11418004590440954, -0.0681122196181562, 0.12861922873994244,
                                           2.104712629544537, 0.08566339835505947, -0.2934078386012377], rtol=1e-11)
    np.testing.assert_allclose(ftl[0][-1], [0.02617631672851312, 0.0164356487254638, -0.11706173095628596,
                                            -1.0325642945215674, 0.3008903460875305, 0.46930183034974804], rtol=1e-11)


def test_ftl_zero_for_constant_curvature():
    """test/test_integration.py:79-90."""
    m = orc.make_model(["constant"], [1], [0.05], [60], 0.5)
    for i in range(-5, 11, 4):
        for j in range(-2, 2):
            prev = orc.solve_g(m, [60 + i * 2], [j], full=False)
            *_, ftl = orc.solve_integrate(m, [0], [i], [j], [30], prev)
            mags = [orc.lib().orc_ftl_magnitude(orc._dp(np.ascontiguousarray(x)), 0) for x in np.concatenate(ftl, 0)]
            # the reference asserts 14 places (LU inverse); the closed-form SE(3) inverse differs by rounding
            assert abs(np.mean(mags)) < 1e-13
    prev = orc.solve_g(m, [64], [-85], full=False)
    *_, ftl = orc.solve_integrate(m, [0.85], [2], [0], [30], prev)
    mags = [orc.lib().orc_ftl_magnitude(orc._dp(np.ascontiguousarray(x)), 0) for x in np.concatenate(ftl, 0)]
    assert abs(np.mean(mags) - 0.85) < 1e-7


def test_soapwg_golden(golden_kin):
    for w, o, g, want in golden_kin["soapwg"]:
        assert orc.lib().orc_soapwg_own_cost(w, o, g) == pytest.approx(want, rel=1e-15)
