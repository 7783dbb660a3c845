"""Independent NumPy cross-check of the oracle's cylinder<->triangle distance.

TEST INFRASTRUCTURE ONLY.  Deliberately shares no case analysis with
oracle/ctr_oracle.c: the distance from a point to the solid cylinder is a
convex function, so its minimum over the triangle is found by a nested
golden-section search over the barycentric parameters (the partial minimum of
a convex function is convex).  Accuracy ~1e-9 mm; used only to arbitrate the
analytic restatement, because python-fcl itself is not installable offline.
"""
import numpy as np

_G = (np.sqrt(5.0) - 1.0) / 2.0


def point_cylinder(x, c, a, h, r):
    """x: (...,3); c,a: (...,3); h,r: (...)"""
    y = x - c
    t = np.sum(y * a, -1)
    w = y - t[..., None] * a
    rho = np.linalg.norm(w, axis=-1)
    dt = np.maximum(np.abs(t) - h, 0.0)
    dr = np.maximum(rho - r, 0.0)
    return np.sqrt(dt * dt + dr * dr)


def _golden(f, lo, hi, iters):
    """vectorised golden-section minimum of f on [lo,hi] (arrays); returns min value seen"""
    x1 = hi - _G * (hi - lo)
    x2 = lo + _G * (hi - lo)
    f1, f2 = f(x1), f(x2)
    best = np.minimum(np.minimum(f(lo), f(hi)), np.minimum(f1, f2))
    for _ in range(iters):
        left = f1 < f2
        hi = np.where(left, x2, hi)
        lo = np.where(left, lo, x1)
        nx1 = hi - _G * (hi - lo)
        nx2 = lo + _G * (hi - lo)
        # re-evaluate both (simple, fully vectorised)
        x1, x2 = nx1, nx2
        f1, f2 = f(x1), f(x2)
        best = np.minimum(best, np.minimum(f1, f2))
    return best


def cylinder_triangle(c, a, h, r, t0, t1, t2, iters=48):
    """all inputs batched on the leading axis P; returns (P,) distances (0 = touching)"""
    c, a, t0, t1, t2 = (np.asarray(v, float) for v in (c, a, t0, t1, t2))
    h = np.asarray(h, float)
    r = np.asarray(r, float)
    e1, e2 = t1 - t0, t2 - t0

    def inner(u):
        def fv(v):
            x = t0 + u[..., None] * e1 + v[..., None] * e2
            return point_cylinder(x, c, a, h, r)
        return _golden(fv, np.zeros_like(u), 1.0 - u, iters)

    return _golden(inner, np.zeros(c.shape[0]), np.ones(c.shape[0]), iters)
