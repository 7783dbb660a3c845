"""Generate golden vectors by importing the UNMODIFIED Python reference.

Run in the development container only (the reference is mounted read-only at
/root/reference and does not exist on the GPU box):

    python tests/golden/make_golden.py

Writes tests/golden/kinematic.npz (+ static.npz when the static part is
enabled).  The reference modules that import here are ctrdapp.model.*,
ctrdapp.heuristic.* and ctrdapp.config.parse_config; ctrdapp.collision needs
python-fcl, which is not installable offline, so there are no FCL-generated
vectors (see DESIGN.md "parity unpinned").
"""
import pathlib
import sys

import numpy as np

REF = pathlib.Path("/root/reference")
sys.path.insert(0, str(REF))

from ctrdapp.model.kinematic import Kinematic, calculate_indices  # noqa: E402
from ctrdapp.model import matrix_utils  # noqa: E402
from ctrdapp.model.strain_bases import get_strains  # noqa: E402
from ctrdapp.heuristic.follow_the_leader import FollowTheLeader, FollowTheLeaderTranslation  # noqa: E402
from ctrdapp.heuristic.obstacles_and_goal import SquareObstacleAvgPlusWeightedGoal  # noqa: E402

HERE = pathlib.Path(__file__).resolve().parent

sys.path.insert(0, str(HERE.parent))
from models_def import MODELS as _MODELS  # noqa: E402

MODELS = [(k,) + tuple(v) for k, v in _MODELS.items()]


def make_model(strain_bases, q_dof, q, lengths, dx):
    qs, o = [], 0
    for d in q_dof:
        qs.append(np.asarray(q[o:o + d], float))
        o += d
    return Kinematic(len(lengths), qs, q_dof, lengths, dx, get_strains(strain_bases, q_dof))


def flat_g(g):
    return np.concatenate([np.asarray(t).reshape(-1, 4, 4) for t in g], 0), np.array([len(t) for t in g])


def main():
    rng = np.random.default_rng(20261019)
    out = {}
    names = []
    for name, bases, dof, q, lengths, dx in MODELS:
        names.append(name)
        m = make_model(bases, dof, q, lengths, dx)
        T = m.tube_num
        N = m.num_discrete_points
        out[f"{name}/npts"] = np.array(N)
        cases_idx, cases_th, cases_full, gs, cnts = [], [], [], [], []
        for c in range(6):
            idx = [int(rng.integers(0, n)) for n in N]
            if c == 0:
                idx = [0] * T
            if c == 1:
                idx = [n - 1 for n in N]
            th = rng.uniform(-2 * np.pi, 2 * np.pi, T)
            if c == 2:
                th[0] = 0.0  # exact theta == 0 branch of exponential_map
            for full in (False, True):
                g, cnt = flat_g(m.solve_g(indices=idx, thetas=list(th), full=full))
                cases_idx.append(idx)
                cases_th.append(th)
                cases_full.append(full)
                gs.append(g)
                cnts.append(cnt)
        out[f"{name}/g_idx"] = np.array(cases_idx)
        out[f"{name}/g_theta"] = np.array(cases_th)
        out[f"{name}/g_full"] = np.array(cases_full)
        out[f"{name}/g_counts"] = np.array(cnts)
        out[f"{name}/g_flat"] = np.concatenate(gs, 0)

        # solve_integrate: parent = solve_g(full=False) at a random config, child = small step
        si = []
        for c in range(4):
            dx_ = m.delta_x
            lengths_f = [float(x) for x in lengths]
            mins = [0.0] + lengths_f[:-1]
            p_idx = [int(rng.integers(round(mins[n] / dx_), N[n])) for n in range(T)]
            p_th = rng.uniform(0, 2 * np.pi, T)
            prev_g = m.solve_g(indices=p_idx, thetas=list(p_th), full=False)
            prev_ins = [lengths_f[n] - p_idx[n] * dx_ for n in range(T)]  # intuitive insertion
            d_ins = rng.uniform(-3 * dx_, 3 * dx_, T)
            d_th = rng.uniform(-0.17, 0.17, T)
            this_ins = [prev_ins[n] + d_ins[n] for n in range(T)]
            this_th = p_th + d_th
            g_out, eta, nidx, tins, ftl = m.solve_integrate(list(d_th), list(d_ins), list(this_th), this_ins, prev_g)
            pg, pc = flat_g(prev_g)
            og, oc = flat_g(g_out)
            key = f"{name}/si{c}"
            out[key + "/prev_g"] = pg
            out[key + "/prev_counts"] = pc
            out[key + "/args"] = np.array([d_th, d_ins, this_th, this_ins])
            out[key + "/g"] = og
            out[key + "/g_counts"] = oc
            out[key + "/eta"] = np.array([e[0] for e in eta])
            out[key + "/new_idx"] = np.array(nidx)
            out[key + "/true_ins"] = np.array(tins)
            out[key + "/ftl"] = np.concatenate([np.asarray(t).reshape(-1, 6) for t in ftl], 0)
            out[key + "/ftl_counts"] = np.array([len(t) for t in ftl])
            out[key + "/ftl_mag"] = np.array([FollowTheLeader(True, ftl).get_cost(),
                                              FollowTheLeader(False, ftl).get_cost(),
                                              FollowTheLeaderTranslation(True, ftl).get_cost(),
                                              FollowTheLeaderTranslation(False, ftl).get_cost()])
            si.append(c)
    out["models"] = np.array(names)

    # calculate_indices: random + near-half cases (kinematic.py:201-274)
    ci_in, ci_out = [], []
    for _ in range(400):
        dx = float(rng.choice([1.0, 0.5, 0.1, 2.5]))
        L = [50.0, 80.0]
        prev = rng.uniform(-5, 90, 2)
        kind = rng.integers(0, 3)
        if kind == 0:
            nxt = prev + rng.uniform(-0.6 * dx, 0.6 * dx, 2)
        elif kind == 1:
            nxt = prev + rng.uniform(-4 * dx, 4 * dx, 2)
        else:
            prev = np.round(prev / dx) * dx + rng.choice([0.0, 0.5 * dx, -0.5 * dx], 2)
            nxt = prev + rng.choice([0.0, 0.5 * dx, -0.5 * dx, dx, 1.5 * dx], 2)
        p, n = calculate_indices(list(prev), list(nxt), L, dx)
        ci_in.append(np.concatenate([prev, nxt, L, [dx]]))
        ci_out.append(np.array(p + n))
    out["calc_idx/in"] = np.array(ci_in)
    out["calc_idx/out"] = np.array(ci_out)

    # exponential_map (matrix_utils.py:157-177)
    em_in, em_out = [], []
    for _ in range(40):
        s = rng.normal(size=6) * rng.choice([1e-6, 1e-3, 0.05, 1.0])
        if rng.random() < 0.2:
            s[0:3] = 0.0
        gh = matrix_utils.hat(s)
        th = float(np.linalg.norm(s[0:3]))
        em_in.append(np.concatenate([[th], gh.ravel()]))
        em_out.append(matrix_utils.exponential_map(th, gh).ravel())
    out["expmap/in"] = np.array(em_in)
    out["expmap/out"] = np.array(em_out)

    # SOAPWG arithmetic (obstacles_and_goal.py)
    sg = []
    for _ in range(32):
        w, o, g = rng.uniform(0.1, 5), rng.uniform(1e-3, 20), rng.uniform(0, 200)
        sg.append([w, o, g, SquareObstacleAvgPlusWeightedGoal(w, o, g).get_cost()])
    out["soapwg"] = np.array(sg)

    np.savez_compressed(HERE / "kinematic.npz", **out)
    print("wrote", HERE / "kinematic.npz", sum(v.nbytes for v in out.values()) / 1e6, "MB raw")


if __name__ == "__main__":
    main()
