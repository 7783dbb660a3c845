"""CPU: the C restatement of the static model (residual, integrate_g, MINPACK hybrd) vs golden
vectors produced by the Python reference (tests/golden/make_golden_static.py)."""
import numpy as np
import pytest

from conftest import GOLDEN
from models_def import STATIC_MODELS
from oracle import oracle as orc


@pytest.fixture(scope="module")
def gs():
    return np.load(GOLDEN / "static.npz")


def _model(name):
    bases, dof, q, lengths, dx, r_out, r_in, degree, cases = STATIC_MODELS[name]
    return orc.make_static(bases, dof, q, lengths, dx, r_out, r_in, degree), cases


@pytest.mark.parametrize("name", list(STATIC_MODELS))
def test_init_guess_and_residual(gs, name):
    s, cases = _model(name)
    np.testing.assert_array_equal(orc.static_init_guess(s), gs[f"{name}/x0"])  # static_bases.py:142-168
    for c, (idx, th) in enumerate(cases):
        for x, F in zip(gs[f"{name}/c{c}/res_x"], gs[f"{name}/c{c}/res_F"]):
            got = orc.static_residual(s, idx, th, x)
            scale = max(np.max(np.abs(F)), 1e-300)
            assert np.max(np.abs(got - F)) / scale < 1e-12


@pytest.mark.parametrize("name", list(STATIC_MODELS))
def test_integrate_g_from_reference_solution(gs, name):
    """Same solution q => same curves (integrate_g with its absolute-index quirk, static.py:165-198)."""
    s, cases = _model(name)
    for c, (idx, th) in enumerate(cases):
        for tag in ("default", "tight"):
            g = orc.static_curves(s, idx, th, gs[f"{name}/c{c}/sol_{tag}"])
            assert [len(t) for t in g] == gs[f"{name}/c{c}/counts_{tag}"].tolist()
            ref = gs[f"{name}/c{c}/g_{tag}"]
            got = np.concatenate(g, 0)
            scale = np.maximum(1.0, np.linalg.norm(ref[:, 0:3, 3], axis=1))[:, None, None]
            # one-step sections evaluate the fitted x^2 strain coefficients at x ~ 119 (the
            # absolute-index quirk), with +-3.5e3 cancelling coefficients: rounding is amplified ~1e10; everywhere else 1e-10 holds
            short = min(s.m.npts[n] - idx[n] for n in range(s.m.tube_num)) <= 2
            assert np.max(np.abs(got - ref) / scale) < (1e-4 if short else 1e-10)


@pytest.mark.parametrize("name", list(STATIC_MODELS))
def test_hybrd_restatement(gs, name):
    """Converged cases reach the reference's root (both tolerances); the evaluation count stays within
    a few of SciPy's (rounding in the residual can shift one Broyden step); the non-converging
    configuration (c3 case 3: 769 evaluations, success False) is reported as not converged too."""
    s, cases = _model(name)
    for c, (idx, th) in enumerate(cases):
        for tag, xtol in (("default", 0.0), ("tight", 1e-13)):
            nfev_ref, ok_ref, _ = gs[f"{name}/c{c}/meta_{tag}"]
            g, sol, nfev, info = orc.static_solve_g(s, idx, th, xtol=xtol)
            assert (info == 1) == bool(ok_ref), (name, c, tag, info)
            if not ok_ref:
                continue
            if min(s.m.npts[n] - idx[n] for n in range(s.m.tube_num)) <= 2:
                continue  # one-step sections: the system is degenerate, roots are not unique
            ref = gs[f"{name}/c{c}/sol_tight"]
            tol = 5e-6 if tag == "default" else 1e-9
            assert np.max(np.abs(sol - ref)) <= tol * max(1.0, np.max(np.abs(ref))), (name, c, tag)
            # SciPy reports two evaluations more than MINPACK's own count (its shape check); the restated
            # iteration follows the same path except where rounding flips a trust-region decision
            if tag == "default":
                assert abs(nfev + 2 - nfev_ref) <= 0.4 * nfev_ref, (name, c, tag, nfev, nfev_ref)
