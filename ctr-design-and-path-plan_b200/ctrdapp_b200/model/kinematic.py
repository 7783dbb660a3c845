"""Kinematic strain-based model over libctrb200 (mirror of ctrdapp/model/kinematic.py).

Reference-shaped methods (B = 1, list-of-lists of 4x4 arrays):
    solve_g, solve_integrate, solve_eta              kinematic.py:53-198
Batch methods (what solve/ and optimize/ call with stacked candidates):
    solve_g_batch, solve_integrate_batch, eta_ftl_batch, calculate_indices_batch
"""
import ctypes as C

import numpy as np

from .. import _lib as L
from .model import Model
from .strain_bases import get_strains


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Kinematic(Model):
    """Same constructor as the reference (kinematic.py:46-51); `strain_base` holds base NAMES."""

    def __init__(self, tube_num, q, q_dof, max_tube_length, delta_x, strain_base, ctx=None):
        super().__init__(tube_num, max_tube_length, delta_x)
        self.strain_base = get_strains(strain_base, q_dof)
        self.strain_bias = np.array([0, 0, 0, 1, 0, 0])
        self.q_dof = q_dof
        self.q = q
        self.ctx = ctx or L.default_context()
        d = self._make_desc()
        self._h = C.c_void_p()
        L.check(L.lib().ctrb_model_create(self.ctx.handle, C.byref(d), C.byref(self._h)))
        npts = (C.c_int32 * L.MAX_TUBES)()
        L.check(L.lib().ctrb_model_num_points(self._h, npts))
        assert list(npts[:tube_num]) == self.num_discrete_points
        self._after_create()

    def _make_desc(self):
        d = L.ModelDesc()
        d.model_type = 0
        d.tube_num = self.tube_num
        for n in range(self.tube_num):
            d.base_kind[n] = L.BASE_KINDS[self.strain_base[n]]
            d.q_dof[n] = self.q_dof[n]
            for k in range(self.q_dof[n]):
                d.q[n][k] = float(self.q[n][k])
            d.tube_length[n] = float(self.max_tube_length[n])
        d.delta_x = float(self.delta_x)
        return d

    def _after_create(self):
        pass

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                L.lib().ctrb_model_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    # ------------------------------------------------------------------ batch API (host numpy buffers)
    def node_offsets(self, idx, full=False):
        idx = _i32(idx).reshape(-1, self.tube_num)
        off = np.zeros(idx.size + 1, np.int64)
        L.check(L.lib().ctrb_g_node_offsets(self.ctx.handle, self._h, idx.shape[0], L.ptr(idx), int(bool(full)),
                                            L.ptr(off), L.MEM_HOST))
        return off

    def solve_g_batch(self, idx, theta, full=False, precision=64):
        """-> (g [total_nodes,4,4], node_offsets [B*T+1])."""
        idx = _i32(idx).reshape(-1, self.tube_num)
        theta = _f64(theta).reshape(-1, self.tube_num)
        off = self.node_offsets(idx, full)
        g = np.empty((int(off[-1]), 4, 4), np.float64 if precision == 64 else np.float32)
        L.check(L.lib().ctrb_solve_g_batch(self.ctx.handle, self._h, idx.shape[0], L.ptr(idx), L.ptr(theta),
                                           int(bool(full)), precision, L.ptr(off), L.ptr(g), L.MEM_HOST))
        return g, off

    def calculate_indices_batch(self, prev_ins, next_ins):
        prev_ins = _f64(prev_ins).reshape(-1, self.tube_num)
        next_ins = _f64(next_ins).reshape(-1, self.tube_num)
        p = np.empty(prev_ins.shape, np.int32)
        q = np.empty(prev_ins.shape, np.int32)
        L.check(L.lib().ctrb_calculate_indices_batch(self.ctx.handle, self._h, prev_ins.shape[0], L.ptr(prev_ins),
                                                     L.ptr(next_ins), L.ptr(p), L.ptr(q), L.MEM_HOST))
        return p, q

    def eta_ftl_batch(self, prev_idx, new_idx, delta_theta, prev_g, prev_offsets, want_ftl=True, velocity=None):
        """-> (eta [B,T,6], ftl [total_nodes,6] or None, ftl_mag [B,4]).  Give new_idx (velocity is then
        (prev_idx - new_idx) * delta_x, kinematic.py:71) or an explicit velocity array."""
        prev_idx = _i32(prev_idx).reshape(-1, self.tube_num)
        new_idx = None if velocity is not None else _i32(new_idx).reshape(-1, self.tube_num)
        velocity = None if velocity is None else _f64(velocity).reshape(-1, self.tube_num)
        delta_theta = _f64(delta_theta).reshape(-1, self.tube_num)
        prev_g = _f64(prev_g).reshape(-1, 4, 4)
        prev_offsets = np.ascontiguousarray(prev_offsets, np.int64)
        B = prev_idx.shape[0]
        eta = np.empty((B, self.tube_num, 6))
        ftl = np.zeros((int(prev_offsets[-1]), 6)) if want_ftl else None  # rows beyond npts - prev_idx stay zero
        mag = np.empty((B, 4))
        L.check(L.lib().ctrb_eta_ftl_batch(self.ctx.handle, self._h, B, L.ptr(prev_idx), L.ptr(new_idx),
                                           L.ptr(velocity), L.ptr(delta_theta), L.ptr(prev_g), L.ptr(prev_offsets), L.ptr(eta),
                                           L.ptr(ftl), L.ptr(mag), L.MEM_HOST))
        return eta, ftl, mag

    def solve_integrate_batch(self, delta_theta, delta_insertion, this_theta, this_insertion, prev_g, prev_offsets,
                              invert_insert=True, need_g_out=True, want_ftl=False):
        """kinematic.py:53-80 for K stacked candidates (what RRT._single_iteration does K times).
        All arguments [K, T]; prev_g / prev_offsets are the parents' curves packed like solve_g_batch's output.
        -> dict(g, offsets, eta [K,T,6], insert_indices [K,T], prev_indices [K,T], true_insertions [K,T],
                ftl (packed, or None), ftl_mag [K,4])"""
        T = self.tube_num
        this_insertion = _f64(this_insertion).reshape(-1, T)
        delta_insertion = _f64(delta_insertion).reshape(-1, T)
        lengths = _f64(self.max_tube_length)[None, :]
        if invert_insert:  # :57-59
            this_insertion = lengths - this_insertion
            delta_insertion = -delta_insertion
        prev_idx, new_idx = self.calculate_indices_batch(this_insertion - delta_insertion, this_insertion)
        g, off = self.solve_g_batch(new_idx, this_theta, full=False) if need_g_out else (None, None)
        eta, ftl, mag = self.eta_ftl_batch(prev_idx, new_idx, delta_theta, prev_g, prev_offsets, want_ftl=want_ftl)
        # kinematic.py:74-78: `length - ind * delta_x` is Python arithmetic -- integer tube lengths and delta_x give integer
        # insertions (tree.txt then reads "3, 6, 9", and TreeFromFile parses them back with int())
        integral = isinstance(self.delta_x, (int, np.integer)) and \
            all(isinstance(v, (int, np.integer)) for v in self.max_tube_length)
        if integral:
            li = np.asarray(self.max_tube_length, np.int64)[None, :]
            true_ins = li - new_idx.astype(np.int64) * int(self.delta_x) if invert_insert else new_idx.astype(np.int64) * int(self.delta_x)
        else:
            true_ins = lengths - new_idx * self.delta_x if invert_insert else new_idx * float(self.delta_x)
        return {"g": g, "offsets": off, "eta": eta, "insert_indices": new_idx, "prev_indices": prev_idx,
                "true_insertions": true_ins, "ftl": ftl, "ftl_mag": mag}

    # ------------------------------------------------------------------ reference-shaped API (B = 1)
    @staticmethod
    def _split(flat, off, T, width):
        return [[flat[k] for k in range(int(off[n]), int(off[n + 1]))] for n in range(T)]

    def solve_g(self, indices=None, thetas=None, full=True):
        """kinematic.py:82-130."""
        if indices is None:
            indices = [disc - 1 for disc in self.num_discrete_points]
        if thetas is None:
            thetas = [0] * self.tube_num
        g, off = self.solve_g_batch([indices], [thetas], full=full)
        return self._split(g, off, self.tube_num, (4, 4))

    def solve_eta(self, velocity_list, prev_insert_indices_list, delta_theta_list, prev_g, curr_g):
        """kinematic.py:132-198."""
        counts = [len(t) for t in prev_g]
        off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        flat = np.concatenate([np.asarray(t, np.float64).reshape(-1, 4, 4) for t in prev_g if len(t)], 0)
        eta, ftl, _ = self.eta_ftl_batch([prev_insert_indices_list], None, [delta_theta_list], flat, off,
                                         velocity=[velocity_list])
        eta_out = [[eta[0, n]] for n in range(self.tube_num)]
        ftl_out = []
        for n in range(self.tube_num):
            need = self.num_discrete_points[n] - prev_insert_indices_list[n]
            ftl_out.append([ftl[int(off[n]) + k] for k in range(need)])
        return eta_out, ftl_out

    def solve_integrate(self, delta_theta, delta_insertion, this_theta, this_insertion, prev_g, invert_insert=True,
                        need_g_out=True):
        """kinematic.py:53-80."""
        if invert_insert:  # :57-59
            this_insertion = [length - ins for ins, length in zip(this_insertion, self.max_tube_length)]
            delta_insertion = [-d for d in delta_insertion]
        prev_insertion = [ins - d for ins, d in zip(this_insertion, delta_insertion)]
        p, q = self.calculate_indices_batch([prev_insertion], [this_insertion])
        prev_idx, new_idx = p[0].tolist(), q[0].tolist()
        g_out = self.solve_g(indices=new_idx, thetas=this_theta, full=False) if need_g_out else [[]]
        velocity = [(pi - ni) * self.delta_x for ni, pi in zip(new_idx, prev_idx)]
        eta_out, ftl_out = self.solve_eta(velocity, prev_idx, delta_theta, prev_g, g_out)
        if invert_insert:
            true_insertions = [length - (ind * self.delta_x) for ind, length in zip(new_idx, self.max_tube_length)]
        else:
            true_insertions = [ind * self.delta_x for ind in new_idx]
        return g_out, eta_out, new_idx, true_insertions, ftl_out


def calculate_indices(prev_insertions, next_insertions, max_tube_lengths, delta_x, ctx=None):
    """Module-level helper with the reference's signature (kinematic.py:201-274); runs on the device
    through a throw-away model handle that only carries the lengths and delta_x."""
    T = len(prev_insertions)
    m = Kinematic(T, [np.zeros(1)] * T, [1] * T, list(max_tube_lengths), delta_x, ["constant"] * T, ctx=ctx)
    p, q = m.calculate_indices_batch([prev_insertions], [next_insertions])
    return p[0].tolist(), q[0].tolist()
