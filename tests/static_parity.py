"""Per-bucket parity of the static hot path (Static.solve_g -> check_collision -> own cost) against the CPU
oracle, shared by tests/test_gpu_headline.py and bench.py's cpu_baseline leg.

hybrd's path is chaotic in its input and the equilibrium is multi-stable: two correct
implementations of the same iteration can land on different roots.  So every configuration is put in one of
three buckets against a reference solve, and each bucket carries its own checked statement:

  A  both converged to the same root (|dq| <= 1e-5 max(1, |q|))
  B  both converged, to different roots
  F  the converged flags differ (or neither converged)

and, independently of the bucket, the curves + collision + cost kernels are held to the collision contract on
the unknowns the GPU itself found (the oracle integrates and queries THOSE unknowns).
"""
import numpy as np

BAND = 1e-6          # mm: boundary band of the collision bit, tolerance of the distances (north star)
ROOT_RTOL = 1e-5     # same-root threshold on the unknowns


def buckets(sol_a, info_a, sol_b, info_b):
    """-> dict of boolean masks A, B, F (see module docstring) for solves a, b of the same configurations."""
    ok_a, ok_b = info_a == 1, info_b == 1
    both = ok_a & ok_b
    scale = np.maximum(1.0, np.max(np.abs(sol_b), axis=1))
    with np.errstate(invalid="ignore"):
        same = np.max(np.abs(sol_a - sol_b), axis=1) <= ROOT_RTOL * scale
    return {"A": both & same, "B": both & ~same, "F": ok_a != ok_b, "N": ~ok_a & ~ok_b}


def one_step(idx, npts):
    """configurations with a one-step section 1 or 2 (exposure of tube 0 .. T-2 equal to one node)"""
    e = np.asarray(npts)[None, :] - 1 - np.asarray(idx)
    return np.any(e[:, :-1] == 1, axis=1)


def rates(mask_dict, sel):
    n = max(int(sel.sum()), 1)
    return {k: float((v & sel).sum()) / n for k, v in mask_dict.items()}


def collision_contract(obs, goal, cost, o_obs, o_goal, o_cost):
    """The collision contract on identical curves: bit exact outside the 1e-6 mm band, distances within 1e-6 mm,
    cost within 1e-9 relative.  -> (violations mask, number of configurations inside the band)."""
    in_band = (np.abs(o_obs) < BAND) & (o_obs > 0)
    bit, o_bit = obs < 0, o_obs < 0
    bad = (bit != o_bit) & ~in_band
    free = ~bit & ~o_bit
    with np.errstate(invalid="ignore"):
        bad |= free & ~(np.abs(obs - o_obs) < BAND)
        bad |= bit & (obs != -1.0)
        gb, o_gb = goal < 0, o_goal < 0
        bad |= gb != o_gb
        bad |= ~gb & ~o_gb & ~(np.abs(goal - o_goal) < BAND)
        if cost is not None and o_cost is not None:
            fin = free & np.isfinite(o_cost)
            bad |= fin & ~(np.abs(cost - o_cost) <= 1e-9 * np.abs(o_cost) + 1e-12)
    return bad, int(in_band.sum())


def residual_norms(orc, s, idx, theta, q, nthreads=None):
    """|F(q)| under the ORACLE's marched residual (static.py:228-454 restated node by node)"""
    r = orc.static_batch(s, idx, theta, q_in=q, want_residual=True, nthreads=nthreads)
    return np.linalg.norm(r["residual"], axis=1)
