"""Shared helpers of the parity tests: build the same model for the CUDA path and the oracle."""
import numpy as np

from models_def import MODELS


def split_q(q, q_dof):
    out, o = [], 0
    for d in q_dof:
        out.append(np.asarray(q[o:o + d], float))
        o += d
    return out


def gpu_model(name):
    from ctrdapp_b200.model.kinematic import Kinematic
    bases, dof, q, lengths, dx = MODELS[name]
    return Kinematic(len(lengths), split_q(q, dof), dof, lengths, dx, bases)


def orc_model(name):
    from oracle import oracle as orc
    bases, dof, q, lengths, dx = MODELS[name]
    return orc.make_model(bases, dof, q, lengths, dx)


def rel_err(got, ref):
    """SURVEY 8c parity contract: max |d| / max(1, ||p||) over the 12 R/p entries of each node."""
    got = np.asarray(got, np.float64).reshape(-1, 4, 4)
    ref = np.asarray(ref, np.float64).reshape(-1, 4, 4)
    scale = np.maximum(1.0, np.linalg.norm(ref[:, 0:3, 3], axis=1))[:, None, None]
    return float(np.max(np.abs(got - ref) / scale)) if len(ref) else 0.0


def random_configs(npts, B, rng, lo_exposure=None, hi_exposure=None):
    """idx[b,t] uniform over valid indices (or bounded exposure e: idx = N-1-e), theta ~ U[0, 2pi)."""
    T = len(npts)
    if lo_exposure is None:
        idx = np.stack([rng.integers(0, npts[t], B) for t in range(T)], 1)
    else:
        e = np.stack([rng.integers(lo_exposure, min(hi_exposure, npts[t] - 1) + 1, B) for t in range(T)], 1)
        idx = np.asarray(npts)[None, :] - 1 - e
    theta = rng.uniform(0, 2 * np.pi, (B, T))
    return idx.astype(np.int32), theta


def gpu_static(name):
    from ctrdapp_b200.model.static import Static
    from models_def import STATIC_MODELS
    bases, dof, q, lengths, dx, r_out, r_in, degree, cases = STATIC_MODELS[name]
    return Static(len(lengths), split_q(q, dof), dof, lengths, dx, bases, {"outer": r_out, "inner": r_in}, degree)


def orc_static(name):
    from oracle import oracle as orc
    from models_def import STATIC_MODELS
    bases, dof, q, lengths, dx, r_out, r_in, degree, cases = STATIC_MODELS[name]
    return orc.make_static(bases, dof, q, lengths, dx, r_out, r_in, degree)
