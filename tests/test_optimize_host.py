"""CPU: the vertex-parallel Nelder-Mead driver against scipy.optimize.minimize(method='Nelder-Mead') -- the routine
the reference calls at neldermead.py:83-88 -- on deterministic objectives, the initial simplex against the formula of the reference's
`calculate_simplex` (its module cannot be imported here: it pulls in pympnn), and the 2-rank gloo evaluator."""
import os
import subprocess
import sys

import numpy as np
import pytest
from scipy import optimize

from ctrdapp_b200.optimize.neldermead import calculate_simplex
from ctrdapp_b200.optimize.simplex import nelder_mead


def rosen(x):
    x = np.asarray(x)
    return float(np.sum(100.0 * (x[1:] - x[:-1] ** 2) ** 2 + (1 - x[:-1]) ** 2))


def bumpy(x):
    x = np.asarray(x)
    return float(np.sum(x ** 2) + 0.3 * np.sum(np.cos(7 * x)) + abs(x[0] - 0.2))


@pytest.mark.parametrize("f,dim,seed", [(rosen, 2, 0), (rosen, 5, 1), (bumpy, 5, 2), (bumpy, 8, 3)])
@pytest.mark.parametrize("speculate", [True, False, 1, 2])
def test_same_iterates_as_scipy(f, dim, seed, speculate):
    rng = np.random.default_rng(seed)
    simplex = rng.uniform(-1.5, 1.5, (dim + 1, dim))
    ref = optimize.minimize(f, simplex[0], method="Nelder-Mead",
                            options={"xatol": 1e-6, "fatol": 1e-6, "initial_simplex": simplex, "maxfev": 4000})
    got = nelder_mead(lambda xs: [f(x) for x in xs], simplex, xatol=1e-6, fatol=1e-6, maxfev=4000, speculate=speculate)
    assert got.nit == ref.nit and got.nfev == ref.nfev and got.success == ref.success
    np.testing.assert_array_equal(got.x, ref.x)
    assert got.fun == ref.fun
    assert (got.nfev_speculative > 0) == bool(speculate)
    # the speculative points the algorithm asked for + the discarded ones account for every extra evaluation
    calls = []
    got2 = nelder_mead(lambda xs: calls.append(len(xs)) or [f(x) for x in xs], simplex, xatol=1e-6, fatol=1e-6, maxfev=4000,
                       speculate=speculate)
    assert sum(calls) == got2.nfev + got2.nfev_speculative


def test_neldermead_driver_tags_speculative_evaluations(monkeypatch):
    """NelderMead.find_min in one process: no speculation by default (one RRT run per evaluation SciPy would make, as in the
    reference, so `optimize_iterations` bounds the real work); with speculation on, optimize_process lists exactly the sequential
    algorithm's evaluations and the discarded speculative runs are kept apart."""
    from ctrdapp_b200.optimize import neldermead as nm

    class FakeSolver:
        found_solution = True

        def __init__(self, cost):
            self.cost = cost

        def get_best_cost(self):
            return self.cost, 0

    monkeypatch.setattr(nm, "create_model", lambda conf, x, ctx=None: np.asarray(x))
    monkeypatch.setattr(nm, "create_solver", lambda model, hf, cc, conf: FakeSolver(rosen(model)))
    simplex = np.random.default_rng(4).uniform(-1.5, 1.5, (4, 3)).tolist()
    conf = {"tube_number": 1, "optimizer_precision": 1e-3, "tube_radius": {"outer": [0.9]}, "strain_bases": ["linear"],
            "tube_lengths": [50], "q_dof": [3], "initial_simplex": simplex, "optimize_iterations": 120}
    ref = nelder_mead(lambda xs: [rosen(x) for x in xs], simplex, xatol=1e-3, fatol=1e-3, maxfev=120, speculate=False)
    for speculate in (None, 3):
        c = dict(conf)
        if speculate is not None:
            c["speculate"] = speculate
        opt = nm.NelderMead(None, None, None, c)
        res = opt.find_min()
        assert res.num_function_eval == ref.nfev and res.num_iterations == ref.nit
        assert len(res.optimize_process) == ref.nfev                      # what the sequential algorithm evaluated
        assert opt.count == ref.nfev + res.num_speculative_eval             # every RRT run is accounted for
        assert (res.num_speculative_eval == 0) == (speculate is None)
        assert len(res.speculative_results) == res.num_speculative_eval
        assert res.owns_best_solver and res.best_rank == 0
        assert res.best_solver.cost == min(e["cost"] for e in res.optimize_process + res.speculative_results)


def test_budget_is_respected():
    simplex = np.random.default_rng(5).uniform(-1, 1, (6, 5))
    got = nelder_mead(lambda xs: [rosen(x) for x in xs], simplex, maxfev=20)
    assert got.nfev <= 20 + 5 and not got.success
    with pytest.raises(ValueError):
        nelder_mead(lambda xs: [0.0] * len(xs), np.zeros((3, 3)))


def test_calculate_simplex_literals():
    """neldermead.py:143-191 for config.yaml's design space (linear bases, q_dof [2, 3], radii [0.9, 0.6], L 120):
    caps = min(initial guess, strain limit); vertex 0 = caps / 2; vertex i+1 quarters cap i."""
    s = calculate_simplex(["linear", "linear"], [0.9, 0.6], [120, 120], [2, 3], None)
    caps = [0.05 / 0.9, 0.05 / 0.9 / 120, 0.05 / 0.6, -(0.05 / 0.6 / 120), 0.05 / 0.6 / 120]
    assert len(s) == 6 and s[0] == [c / 2 for c in caps]
    for i in range(5):
        want = list(caps)
        want[i] /= 4
        assert s[i + 1] == want
    g = [0.02, 0.001, 0.05, 0.0001, 0.002]  # guesses below a cap win (:171-174)
    s2 = calculate_simplex(["linear", "linear"], [0.9, 0.6], [120, 120], [2, 3], g)
    assert s2[0][0] == 0.01 and s2[0][2] == 0.025 and s2[0][3] == caps[3] / 2


def test_distributed_evaluator_two_ranks(tmp_path):
    """world_size 2 over gloo: both ranks reach the single-process result, each having evaluated about half."""
    worker = os.path.join(os.path.dirname(__file__), "_nm_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29731")
    procs = [subprocess.Popen([sys.executable, worker, str(r), "2", str(tmp_path)], env=env) for r in range(2)]
    assert all(p.wait(timeout=240) == 0 for p in procs)
    simplex = np.random.default_rng(11).uniform(-1.5, 1.5, (6, 5))
    single = nelder_mead(lambda xs: [rosen(x) for x in xs], simplex, xatol=1e-6, fatol=1e-6, maxfev=3000)
    for r in range(2):
        got = np.load(tmp_path / f"rank{r}.npz")
        np.testing.assert_array_equal(got["x"], single.x)
        assert int(got["nit"]) == single.nit and int(got["nfev"]) == single.nfev
        assert 0.35 < int(got["calls"]) / (single.nfev + single.nfev_speculative) < 0.65
