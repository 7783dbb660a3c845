"""Generate tests/golden/static.npz by importing the UNMODIFIED Python reference's static model
(ctrdapp.model.static via create_model).  Dev container only:  python tests/golden/make_golden_static.py

Per model (C1: 2 tubes / 15 unknowns, C3: 3 tubes / 30 unknowns) and configuration it stores
the residual static_equilibrium(x) at the initial guess and at perturbed points, the MINPACK
solution at SciPy's default xtol (what the reference returns) and at xtol=1e-13 (a tightly
converged root, SURVEY 7 hard part 2), nfev / success, and the g-curves the reference returns.
"""
import contextlib
import io
import pathlib
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))

from ctrdapp.model.model_factory import create_model  # noqa: E402
from ctrdapp.model import static as static_mod  # noqa: E402
from scipy import optimize  # noqa: E402
from models_def import STATIC_MODELS  # noqa: E402


def section_indices(m, idx):
    T = m.tube_num
    N = m.num_discrete_points
    i11, i22, i33 = idx[0], idx[T - 2], idx[T - 1]
    i21 = i22 - (N[0] - i11 - 1)
    i32 = i33 - (N[T - 2] - i22 - 1)
    i31 = i32 - (N[0] - i11 - 1)
    if T == 2:
        i11 = i21 = i31 = 0
    return {(1, 1): i11, (2, 1): i21, (3, 1): i31, (2, 2): i22, (3, 2): i32, (3, 3): i33}


def main():
    rng = np.random.default_rng(20261020)
    out = {}
    real_root = optimize.root
    for name, (bases, dof, q, lengths, dx, r_out, r_in, degree, cases) in STATIC_MODELS.items():
        cfg = dict(model_type="static", tube_number=len(lengths), strain_bases=bases, q_dof=dof, tube_lengths=lengths,
                   delta_x=dx, tube_radius={"outer": r_out, "inner": r_in}, degree=degree, basis_type="last_strain_base")
        with contextlib.redirect_stdout(io.StringIO()):
            m = create_model(cfg, q)
        x0 = np.asarray(m.general_init_guess, float)
        out[f"{name}/x0"] = x0
        T = m.tube_num
        for c, (idx, th) in enumerate(cases):
            key = f"{name}/c{c}"
            m.section_indices = section_indices(m, idx)
            m.theta = list(th)
            xs = [x0, x0 + rng.normal(size=len(x0)) * 1e-3, x0 * (1 + 0.1 * rng.normal(size=len(x0)))]
            out[key + "/res_x"] = np.array(xs)
            out[key + "/res_F"] = np.array([m.static_equilibrium(x) for x in xs])
            m.section_indices = None
            res = {}
            for tag, xtol in (("default", None), ("tight", 1e-13)):
                info = {}

                def patched(f, x0, method="hybr", _xtol=xtol, _info=info):
                    sol = real_root(f, x0, method=method) if _xtol is None else \
                        real_root(f, x0, method=method, options={"xtol": _xtol})
                    _info["sol"] = sol
                    return sol
                static_mod.optimize.root = patched
                with contextlib.redirect_stdout(io.StringIO()):
                    g = m.solve_g(indices=list(idx), thetas=list(th))
                static_mod.optimize.root = real_root
                sol = info["sol"]
                res[tag] = sol
                out[key + f"/sol_{tag}"] = np.asarray(sol.x)
                out[key + f"/meta_{tag}"] = np.array([sol.nfev, int(sol.success), float(np.linalg.norm(sol.fun))])
                out[key + f"/g_{tag}"] = np.concatenate([np.asarray(t).reshape(-1, 4, 4) for t in g], 0)
                out[key + f"/counts_{tag}"] = np.array([len(t) for t in g])
            out[key + "/idx"] = np.array(idx)
            out[key + "/theta"] = np.array(th)
            print(name, c, idx, "nfev", res["default"].nfev, res["tight"].nfev, "success", res["default"].success,
                  "|dsol|", np.max(np.abs(res["default"].x - res["tight"].x)), flush=True)
    np.savez_compressed(HERE / "static.npz", **out)
    print("wrote", HERE / "static.npz")


if __name__ == "__main__":
    main()
