"""Static strain-based model over libctrb200 (mirror of ctrdapp/model/static.py).

    Static.solve_g(indices, thetas, initial_guess=None, full=False)      static.py:71-135

The per-configuration MINPACK hybrd solve and the Magnus integration of the section curves run
on the device (csrc/ctrb_static.cu).  Only basis_type 'last_strain_base' with degree 2 exists,
which is also the only combination that runs in the reference.  solve_eta /
solve_integrate raise like the reference's do (static.py:150 adds 4x4 matrices to a 6-vector).
"""
import ctypes as C

import numpy as np

from .. import _lib as L
from .kinematic import Kinematic, _f64, _i32


def check_degree(degree):
    """`degree` of the section polynomials (static_bases.py:61-74).  The reference's _simple_basis writes the three
    curvature components at the hard-coded column offsets 0, 3 and 6 of a 3 (degree + 1)-column matrix, so only degree 2
    gives a well-formed basis (probe: tools/probe_static_degree.py, profiles/r2_probe_static_degree.txt):
      degree 0, 1  IndexError in the reference (columns 3.. / 6.. do not exist)      -> the same exception here
      degree >= 3  the reference runs, on a basis in which two curvature components share a coefficient (columns 3 and 6)
                   and the trailing coefficients multiply nothing (zero Jacobian columns): its output is an artefact of
                   MINPACK's singular-matrix handling, and three tubes need 39 unknowns         -> NotImplementedError"""
    degree = int(degree)
    if degree < 2:
        raise IndexError(f"index {3 if degree < 1 else 6} is out of bounds for axis 1 with size {3 * (degree + 1)} "
                         f"(static degree {degree}: the reference fails the same way, static_bases.py:68)")
    if degree > 2:
        raise NotImplementedError(f"static degree {degree}: the reference's basis is malformed for degree > 2 (static_bases.py:68 "
                                  f"overlaps the curvature components at columns 3 and 6 and leaves {2 * (degree - 2)} coefficients "
                                  f"unused); only degree 2 is built")


class Static(Kinematic):
    def __init__(self, tube_num, q, q_dof, max_tube_length, delta_x, strain_base, radius, degree=2,
                 basis_type="last_strain_base", ctx=None):
        if tube_num not in {3, 2}:
            raise ValueError(f'{tube_num} is not supported by the static model')  # static.py:77-78
        if basis_type != "last_strain_base":
            raise NameError(f'{basis_type} is not a static basis this library builds; the reference\'s "simple" and '
                            f'"all_strain_bases" do not run either (SURVEY App. B#2)')
        check_degree(degree)
        self._static_desc = (radius, degree)
        super().__init__(tube_num, q, q_dof, max_tube_length, delta_x, strain_base, ctx=ctx)

    def _make_desc(self):
        d = super()._make_desc()
        radius, degree = self._static_desc
        d.model_type = 1
        for n in range(self.tube_num):
            d.radius_outer[n] = float(radius.get('outer')[n])
            d.radius_inner[n] = float(radius.get('inner')[n])
        d.static_degree = int(degree)
        d.static_basis_type = 0
        return d

    def _after_create(self):
        nq = C.c_int32()
        L.check(L.lib().ctrb_static_num_unknowns(self._h, C.byref(nq)))
        self.num_unknowns = nq.value
        x0 = np.zeros(self.num_unknowns)
        L.check(L.lib().ctrb_static_initial_guess(self._h, L.ptr(x0)))
        self.general_init_guess = x0  # static_bases.py:142-168

    # ------------------------------------------------------------------ batch
    def static_equilibrium_batch(self, idx, theta, q):
        """Residual F(q) of static.py:228-259 for B (configuration, q) pairs -> [B, nq]."""
        idx = _i32(idx).reshape(-1, self.tube_num)
        theta = _f64(theta).reshape(-1, self.tube_num)
        q = _f64(q).reshape(-1, self.num_unknowns)
        F = np.empty_like(q)
        L.check(L.lib().ctrb_static_residual_batch(self.ctx.handle, self._h, idx.shape[0], L.ptr(idx), L.ptr(theta),
                                                   L.ptr(q), L.ptr(F), L.MEM_HOST))
        return F

    def solve_g_batch(self, idx, theta, initial_guess=None, xtol=0.0, want_g=True, **_):
        """-> dict(g [nodes,4,4], offsets, solution [B,nq], nfev [B], info [B], converged [B])."""
        idx = _i32(idx).reshape(-1, self.tube_num)
        theta = _f64(theta).reshape(-1, self.tube_num)
        B = idx.shape[0]
        off = self.node_offsets(idx, False) if want_g else None
        g = np.empty((int(off[-1]), 4, 4)) if want_g else None
        sol = np.empty((B, self.num_unknowns))
        nfev = np.empty(B, np.int32)
        info = np.empty(B, np.int32)
        x0, per = None, 0
        if initial_guess is not None:
            x0 = _f64(initial_guess)
            per = int(x0.ndim == 2)
        L.check(L.lib().ctrb_static_solve_g_batch(self.ctx.handle, self._h, B, L.ptr(idx), L.ptr(theta), L.ptr(x0), per,
                                                  float(xtol), L.ptr(off), L.ptr(g), L.ptr(sol), L.ptr(nfev),
                                                  L.ptr(info), L.MEM_HOST))
        return {"g": g, "offsets": off, "solution": sol, "nfev": nfev, "info": info, "converged": info == 1}

    def curves_batch(self, idx, theta, q):
        """integrate_g alone (static.py:165-198 through the calls of :110-130): the section curves of GIVEN
        unknowns q [B, nq] -> dict(g [nodes,4,4], offsets)."""
        idx = _i32(idx).reshape(-1, self.tube_num)
        theta = _f64(theta).reshape(-1, self.tube_num)
        q = _f64(q).reshape(-1, self.num_unknowns)
        off = self.node_offsets(idx, False)
        g = np.empty((int(off[-1]), 4, 4))
        L.check(L.lib().ctrb_static_curves_batch(self.ctx.handle, self._h, idx.shape[0], L.ptr(idx), L.ptr(theta), L.ptr(q),
                                                 L.ptr(off), L.ptr(g), L.MEM_HOST))
        return {"g": g, "offsets": off}

    # ------------------------------------------------------------------ reference-shaped
    def solve_g(self, indices=None, thetas=None, initial_guess=None, full=False):
        """static.py:71-135 (`full` is ignored there as well).  The reference prints nfev and q per
        solve (static.py:105-107); this mirror stores them in .last_solution instead."""
        if indices is None:
            indices = [disc - 1 for disc in self.num_discrete_points]
        if thetas is None:
            thetas = [0] * self.tube_num
        res = self.solve_g_batch([indices], [thetas], initial_guess=initial_guess)
        self.last_solution = {"x": res["solution"][0], "nfev": int(res["nfev"][0]), "success": bool(res["converged"][0])}
        return self._split(res["g"], res["offsets"], self.tube_num, (4, 4))

    def solve_eta(self, velocity_list, prev_insert_indices_list, delta_theta_list, prev_g, curr_g):
        raise ValueError("Static.solve_eta is not usable in the reference either: static.py:150 adds 4x4 matrices to "
                         "a 6-vector (SURVEY App. B#1)")

    def solve_integrate(self, *args, **kwargs):
        raise ValueError("Static.solve_integrate is unusable in the reference (it calls Static.solve_eta, "
                         "SURVEY App. B#1); use solve_g")
