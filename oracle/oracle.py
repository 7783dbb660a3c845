"""ctypes front end of the CPU oracle (oracle/ctr_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package never
imports this module.
"""
import ctypes as C
import json
import os
import pathlib
import subprocess

import numpy as np

_HERE = pathlib.Path(__file__).resolve().parent
MAX_TUBES = 4
MAX_DOF = 8

BASE_KINDS = {"constant": 0, "linear": 1, "quadratic": 2, "pure_helix": 3, "linear_helix": 4,
              "torsion_helix": 5, "torsion_linear_helix": 6, "full": 7}
SHAPE_NONE, SHAPE_SPHERE, SHAPE_BOX, SHAPE_CYLINDER, SHAPE_MESH = range(5)
DBL_MAX = float(np.finfo(np.float64).max)


class OrcModel(C.Structure):
    _fields_ = [("tube_num", C.c_int),
                ("base_kind", C.c_int * MAX_TUBES),
                ("q_dof", C.c_int * MAX_TUBES),
                ("q", (C.c_double * MAX_DOF) * MAX_TUBES),
                ("length", C.c_double * MAX_TUBES),
                ("delta_x", C.c_double),
                ("npts", C.c_int * MAX_TUBES)]


class OrcPrim(C.Structure):
    _fields_ = [("shape", C.c_int), ("center", C.c_double * 3), ("rot", C.c_double * 9),
                ("dims", C.c_double * 3)]


class OrcStatic(C.Structure):
    _fields_ = [("m", OrcModel), ("r_out", C.c_double * MAX_TUBES), ("r_in", C.c_double * MAX_TUBES),
                ("degree", C.c_int)]


def build(force=False):
    so = _HERE / "libctr_oracle.so"
    srcs = [_HERE / "ctr_oracle.c", _HERE / "ctr_oracle.h", _HERE / "static_oracle.c"]
    newest = max(p.stat().st_mtime for p in srcs if p.exists())
    if force or not so.exists() or so.stat().st_mtime < newest:
        subprocess.run(["make", "-C", str(_HERE), "-B"], check=True, capture_output=True)
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build()))
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int)
        _lib.orc_point_cylinder.restype = C.c_double
        _lib.orc_point_cylinder.argtypes = [dp, dp, dp, C.c_double, C.c_double]
        _lib.orc_segment_cylinder.restype = C.c_double
        _lib.orc_segment_cylinder.argtypes = [dp, dp, dp, dp, C.c_double, C.c_double]
        _lib.orc_cylinder_triangle.restype = C.c_double
        _lib.orc_cylinder_triangle.argtypes = [dp, dp, C.c_double, C.c_double, dp, dp, dp]
        _lib.orc_point_triangle.restype = C.c_double
        _lib.orc_point_triangle.argtypes = [dp, dp, dp, dp]
        _lib.orc_soapwg_own_cost.restype = C.c_double
        _lib.orc_soapwg_own_cost.argtypes = [C.c_double] * 3
        _lib.orc_ftl_magnitude.restype = C.c_double
        _lib.orc_ftl_magnitude.argtypes = [dp, C.c_int]
        _lib.orc_env_create.restype = C.c_void_p
        _lib.orc_env_create.argtypes = [dp, C.c_int, C.POINTER(OrcPrim), C.c_int, C.POINTER(OrcPrim), dp, C.c_int]
        _lib.orc_env_destroy.argtypes = [C.c_void_p]
        _lib.orc_env_check_collision.argtypes = [C.c_void_p, dp, ip, C.c_int, dp, C.c_int, dp]
        _lib.orc_env_eval_batch.argtypes = [C.POINTER(OrcModel), C.c_void_p, C.c_int, ip, dp, dp, C.c_double,
                                            dp, dp, dp, C.c_int]
        _lib.orc_kin_solve_g_batch.argtypes = [C.POINTER(OrcModel), C.c_int, ip, dp, dp, C.c_int]
        _lib.orc_exponential_map.argtypes = [C.c_double, dp, dp]
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def py_round_half_even(x):
    return int(np.rint(x))


def make_model(strain_bases, q_dof, q, tube_lengths, delta_x):
    """Mirror of create_model(...)'s kinematic branch (model_factory.py:12-70)."""
    T = len(tube_lengths)
    m = OrcModel()
    m.tube_num = T
    flat = list(q)
    if len(flat) == sum(q_dof) and not isinstance(flat[0], (list, tuple, np.ndarray)):
        qs, o = [], 0
        for d in q_dof:
            qs.append(flat[o:o + d])
            o += d
    else:
        qs = [list(x) for x in q]
    for n in range(T):
        m.base_kind[n] = BASE_KINDS[strain_bases[n]]
        m.q_dof[n] = q_dof[n]
        for k in range(q_dof[n]):
            m.q[n][k] = float(qs[n][k])
        m.length[n] = float(tube_lengths[n])
        m.npts[n] = round(tube_lengths[n] / delta_x) + 1  # model.py:37-38
    m.delta_x = float(delta_x)
    return m


def exponential_map(theta, gamma_hat):
    out = np.zeros((4, 4))
    lib().orc_exponential_map(float(theta), _dp(_f64(gamma_hat)), _dp(out))
    return out


def calculate_indices(prev_ins, next_ins, max_len, delta_x):
    n = len(prev_ins)
    p = np.zeros(n, np.int32)
    q = np.zeros(n, np.int32)
    lib().orc_calculate_indices(n, _dp(_f64(prev_ins)), _dp(_f64(next_ins)), _dp(_f64(max_len)),
                                C.c_double(delta_x), _ip(p), _ip(q))
    return p.tolist(), q.tolist()


def _split(flat, counts, width):
    out, o = [], 0
    for c in counts:
        out.append(flat[o:o + c].reshape(c, *width).copy())
        o += c
    return out


def solve_g(m, indices=None, thetas=None, full=True):
    """Kinematic.solve_g (kinematic.py:82-130) -> list[T] of (n,4,4) arrays."""
    T = m.tube_num
    if indices is None:
        indices = [m.npts[n] - 1 for n in range(T)]
    if thetas is None:
        thetas = [0.0] * T
    tot = sum(m.npts[n] for n in range(T))
    g = np.zeros((tot, 4, 4))
    counts = np.zeros(T, np.int32)
    idx = np.ascontiguousarray(indices, np.int32)
    lib().orc_kin_solve_g(C.byref(m), _ip(idx), _dp(_f64(thetas)), int(bool(full)), _dp(g), _ip(counts))
    return _split(g, counts.tolist(), (4, 4))


def _pack_g(g_list):
    counts = np.array([len(t) for t in g_list], np.int32)
    flat = np.concatenate([np.asarray(t, np.float64).reshape(-1, 4, 4) for t in g_list if len(t)], 0) \
        if counts.sum() else np.zeros((0, 4, 4))
    return np.ascontiguousarray(flat), counts


def solve_eta(m, velocity, prev_idx, delta_theta, prev_g):
    T = m.tube_num
    flat, counts = _pack_g(prev_g)
    tot = sum(m.npts[n] for n in range(T))
    eta = np.zeros((T, 6))
    ftl = np.zeros((tot, 6))
    fc = np.zeros(T, np.int32)
    rc = lib().orc_kin_solve_eta(C.byref(m), _dp(_f64(velocity)), _ip(np.ascontiguousarray(prev_idx, np.int32)),
                                 _dp(_f64(delta_theta)), _dp(flat), _ip(counts), _dp(eta), _dp(ftl), _ip(fc))
    if rc:
        raise IndexError("prev_g shorter than the previous exposure (reference raises IndexError)")
    return [[e] for e in eta], _split(ftl, fc.tolist(), (6,))


def solve_integrate(m, delta_theta, delta_ins, this_theta, this_ins, prev_g, invert_insert=True, need_g_out=True):
    T = m.tube_num
    flat, counts = _pack_g(prev_g)
    tot = sum(m.npts[n] for n in range(T))
    g = np.zeros((tot, 4, 4))
    gc = np.zeros(T, np.int32)
    eta = np.zeros((T, 6))
    ftl = np.zeros((tot, 6))
    fc = np.zeros(T, np.int32)
    new_idx = np.zeros(T, np.int32)
    true_ins = np.zeros(T)
    rc = lib().orc_kin_solve_integrate(C.byref(m), _dp(_f64(delta_theta)), _dp(_f64(delta_ins)),
                                       _dp(_f64(this_theta)), _dp(_f64(this_ins)), _dp(flat), _ip(counts),
                                       int(invert_insert), int(need_g_out), _dp(g), _ip(gc), _dp(eta),
                                       _ip(new_idx), _dp(true_ins), _dp(ftl), _ip(fc))
    if rc:
        raise IndexError("prev_g shorter than the previous exposure")
    g_out = _split(g, gc.tolist(), (4, 4)) if need_g_out else [[]]
    return g_out, [[e] for e in eta], new_idx.tolist(), true_ins.tolist(), _split(ftl, fc.tolist(), (6,))


# ---------------------------------------------------------------- geometry
def point_cylinder(x, c, a, h, r):
    return lib().orc_point_cylinder(_dp(_f64(x)), _dp(_f64(c)), _dp(_f64(a)), h, r)


def segment_cylinder(p0, p1, c, a, h, r):
    return lib().orc_segment_cylinder(_dp(_f64(p0)), _dp(_f64(p1)), _dp(_f64(c)), _dp(_f64(a)), h, r)


def cylinder_triangle(c, a, h, r, t0, t1, t2):
    return lib().orc_cylinder_triangle(_dp(_f64(c)), _dp(_f64(a)), h, r, _dp(_f64(t0)), _dp(_f64(t1)), _dp(_f64(t2)))


def cylinder_cylinder(c1, a1, h1, r1, c2, a2, h2, r2):
    f = lib().orc_cylinder_cylinder
    f.restype = C.c_double
    f.argtypes = [C.POINTER(C.c_double)] * 2 + [C.c_double] * 2 + [C.POINTER(C.c_double)] * 2 + [C.c_double] * 2
    return f(_dp(_f64(c1)), _dp(_f64(a1)), h1, r1, _dp(_f64(c2)), _dp(_f64(a2)), h2, r2)


def point_triangle(x, t0, t1, t2):
    return lib().orc_point_triangle(_dp(_f64(x)), _dp(_f64(t0)), _dp(_f64(t1)), _dp(_f64(t2)))


# ---------------------------------------------------------------- environment
def read_stl(path):
    """Triangle soup [n,3,3] float32, as numpy-stl's Mesh.points gives it (init_collision.py:59-88).
    Binary detection by size (the colon files' 80-byte header starts with 'solid')."""
    raw = pathlib.Path(path).read_bytes()
    if len(raw) >= 84:
        n = int(np.frombuffer(raw, "<u4", 1, 80)[0])
        if 84 + 50 * n == len(raw):
            rec = np.frombuffer(raw, np.dtype([("n", "<f4", 3), ("v", "<f4", 9), ("a", "<u2")]), n, 84)
            return rec["v"].reshape(n, 3, 3).astype(np.float32)
    verts = [ln.split()[1:4] for ln in raw.decode("ascii", "replace").splitlines()
             if ln.strip().startswith("vertex")]
    return np.asarray(verts, np.float32).reshape(-1, 3, 3)


def rotate_align(vector_to, vector_from=(0.0, 0.0, 1.0)):
    """init_collision.py:149-188."""
    vector_to = np.asarray(vector_to, float)
    vector_from = np.asarray(vector_from, float)
    axis = np.cross(vector_from, vector_to)
    cos_a = np.dot(vector_from, vector_to)
    k = 1 / (1 + cos_a)
    return np.array([[axis[0] * axis[0] * k + cos_a, axis[1] * axis[0] * k - axis[2], axis[2] * axis[0] * k + axis[1]],
                     [axis[0] * axis[1] * k + axis[2], axis[1] * axis[1] * k + cos_a, axis[2] * axis[1] * k - axis[0]],
                     [axis[0] * axis[2] * k - axis[1], axis[1] * axis[2] * k + axis[0], axis[2] * axis[2] * k + cos_a]])


def _prim_from_json(o):
    p = OrcPrim()
    p.rot[:] = np.eye(3).ravel().tolist()
    p.center[:] = [float(x) for x in o.get("transform")]
    shape = o.get("shape")
    if shape == "Sphere":
        p.shape = SHAPE_SPHERE
        p.dims[:] = [float(o.get("radius")), 0.0, 0.0]
    elif shape == "Box":
        p.shape = SHAPE_BOX
        p.dims[:] = [float(o.get("x_length")), float(o.get("y_length")), float(o.get("z_length"))]
    elif shape == "Cylinder":
        p.shape = SHAPE_CYLINDER
        p.dims[:] = [float(o.get("radius")), float(o.get("length")), 0.0]
        p.rot[:] = rotate_align(np.asarray(o.get("direction"))).ravel().tolist()
    else:
        raise NotImplementedError(f"Shape type {shape} not supported.")  # init_collision.py:140
    return p


class Env:
    """Obstacles + goal of one objects-JSON (init_collision.py:14-146) or given arrays."""

    def __init__(self, tris=None, prims=(), goal=None, goal_tris=None):
        self.tris = _f64(tris).reshape(-1, 9) if tris is not None and len(tris) else np.zeros((0, 9))
        self.goal_tris = _f64(goal_tris).reshape(-1, 9) if goal_tris is not None and len(goal_tris) else np.zeros((0, 9))
        self.prims = list(prims)
        self.goal = goal
        arr = (OrcPrim * max(1, len(self.prims)))(*self.prims)
        self._h = lib().orc_env_create(_dp(self.tris), len(self.tris), arr, len(self.prims),
                                       C.byref(goal) if goal is not None else None,
                                       _dp(self.goal_tris), len(self.goal_tris))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_env_destroy(self._h)
            self._h = None

    @classmethod
    def from_json(cls, path):
        path = pathlib.Path(path)
        full = json.loads(path.read_text())

        def mesh_of(o):
            t = read_stl(path.parent / o.get("filename")).astype(np.float64)
            return t + np.asarray(o.get("transform"), np.float64)  # translation only, init_collision.py:125

        tris, prims = [], []
        for o in full.get("obstacles"):
            if o.get("mesh"):
                tris.append(mesh_of(o).reshape(-1, 9))
            else:
                prims.append(_prim_from_json(o))
        goal, goal_tris = None, None
        goals = full.get("goal")
        if goals:
            g0 = goals[0]  # init_collision.py:34-39: first goal only
            if g0.get("mesh"):
                goal = OrcPrim()
                goal.shape = SHAPE_MESH
                goal_tris = mesh_of(g0).reshape(-1, 9)
            else:
                goal = _prim_from_json(g0)
        return cls(np.concatenate(tris, 0) if tris else None, prims, goal, goal_tris)

    def check_collision(self, curve, rad, brute=False):
        """CollisionChecker.check_collision (collision_checker.py:35-70)."""
        counts = np.array([len(t) for t in curve], np.int32)
        pos = np.concatenate([np.asarray(t, np.float64).reshape(-1, 4, 4)[:, 0:3, 3] for t in curve if len(t)], 0)
        pos = np.ascontiguousarray(pos)
        out = np.zeros(2)
        rc = lib().orc_env_check_collision(self._h, _dp(pos), _ip(counts), len(curve), _dp(_f64(rad)),
                                           int(brute), _dp(out))
        if rc:
            raise NotImplementedError("cylinder obstacle vs tube cylinder is not restated in the oracle")
        return float(out[0]), float(out[1])

    def eval_batch(self, m, idx, theta, rad, goal_weight, nthreads=None):
        idx = np.ascontiguousarray(idx, np.int32)
        theta = _f64(theta)
        B = idx.shape[0]
        obs = np.zeros(B)
        goal = np.zeros(B)
        cost = np.zeros(B)
        nthreads = nthreads or len(os.sched_getaffinity(0))
        rc = lib().orc_env_eval_batch(C.byref(m), self._h, B, _ip(idx), _dp(theta), _dp(_f64(rad)),
                                      float(goal_weight), _dp(obs), _dp(goal), _dp(cost), nthreads)
        if rc:
            raise NotImplementedError("unsupported obstacle in oracle")
        return obs, goal, cost


def static_eval_batch(env, s, idx, theta, rad, goal_weight, nthreads=None):
    """Static.solve_g -> check_collision -> own cost for a batch (pthreads)."""
    idx = np.ascontiguousarray(idx, np.int32)
    theta = _f64(theta)
    B = idx.shape[0]
    obs, goal, cost = np.zeros(B), np.zeros(B), np.zeros(B)
    info, nfev = np.zeros(B, np.int32), np.zeros(B, np.int32)
    lib().orc_static_eval_batch(C.byref(s), C.c_void_p(env._h), B, _ip(idx), _dp(theta), _dp(_f64(rad)),
                                C.c_double(goal_weight), _dp(obs), _dp(goal), _dp(cost), _ip(info), _ip(nfev),
                                nthreads or len(os.sched_getaffinity(0)))
    return obs, goal, cost, info, nfev


def solve_g_tips_batch(m, idx, theta, nthreads=None):
    idx = np.ascontiguousarray(idx, np.int32)
    theta = _f64(theta)
    B = idx.shape[0]
    tips = np.zeros((B, 4, 4))
    lib().orc_kin_solve_g_batch(C.byref(m), B, _ip(idx), _dp(theta), _dp(tips), nthreads or len(os.sched_getaffinity(0)))
    return tips


def build_tube(curve, rad):
    """CollisionChecker._build_tube (collision_checker.py:98-148): (radius, length, centre, quaternion wxyz)."""
    out = []
    for n in range(len(curve)):
        for i in range(len(curve[n]) - 1):
            f = np.asarray(curve[n][i])[0:3, 3]
            t = np.asarray(curve[n][i + 1])[0:3, 3]
            vec = t - f
            mid = f + vec / 2
            length = np.linalg.norm(vec)
            u = vec / length
            if u[2] == -1.0:
                quat = np.array([0.0, 1.0, 0.0, 0.0])
            else:
                quat = np.array([1 + u[2], -u[1], u[0], 0.0])
                quat = quat / np.linalg.norm(quat)
            out.append((rad[n], length, mid, quat))
    return out


# ---------------------------------------------------------------- static model
def make_static(strain_bases, q_dof, q, tube_lengths, delta_x, r_out, r_in, degree=2):
    s = OrcStatic()
    s.m = make_model(strain_bases, q_dof, q, tube_lengths, delta_x)
    for n in range(len(tube_lengths)):
        s.r_out[n] = float(r_out[n])
        s.r_in[n] = float(r_in[n])
    s.degree = degree
    return s


def static_init_guess(s):
    x0 = np.zeros(64)
    n = lib().orc_static_init_guess(C.byref(s), _dp(x0))
    return x0[:n].copy()


def static_residual(s, indices, thetas, q):
    F = np.zeros(64)
    n = lib().orc_static_residual(C.byref(s), _ip(np.ascontiguousarray(indices, np.int32)), _dp(_f64(thetas)),
                                  _dp(_f64(q)), _dp(F))
    return F[:n].copy()


def static_curves(s, indices, thetas, sol):
    T = s.m.tube_num
    tot = sum(s.m.npts[n] for n in range(T))
    g = np.zeros((tot, 4, 4))
    counts = np.zeros(T, np.int32)
    lib().orc_static_curves(C.byref(s), _ip(np.ascontiguousarray(indices, np.int32)), _dp(_f64(thetas)),
                            _dp(_f64(sol)), _dp(g), _ip(counts))
    return _split(g, counts.tolist(), (4, 4))


def static_batch(s, idx, theta, env=None, rad=None, goal_weight=0.0, one_step_rule=0, x0=None, q_in=None,
                 want_residual=False, nthreads=None):
    """Threaded batch behind the parity tests (orc_static_batch).  Without q_in: Static.solve_g per configuration
    (one_step_rule 0 = the reference's literal hybrd, 1 = the reduced system for one-step sections) from x0 [B, n] or
    the default guess.  With q_in [B, n]: no solve, the given unknowns are evaluated.  env: also curves -> collision
    -> own cost.  -> dict(solution, info, nfev[, residual][, obs, goal, cost])."""
    idx = np.ascontiguousarray(idx, np.int32)
    theta = _f64(theta)
    B = idx.shape[0]
    n = static_init_guess(s).shape[0]
    sol = np.zeros((B, n))
    fres = np.zeros((B, n)) if want_residual else None
    info, nfev = np.zeros(B, np.int32), np.zeros(B, np.int32)
    obs = goal = cost = None
    if env is not None:
        obs, goal, cost = np.zeros(B), np.zeros(B), np.zeros(B)
    x0a = None if x0 is None else _f64(x0).reshape(B, n)
    qa = None if q_in is None else _f64(q_in).reshape(B, n)
    f = lib().orc_static_batch
    f.argtypes = None
    nul = None
    rc = f(C.byref(s), C.c_void_p(env._h) if env is not None else nul, C.c_int(B), _ip(idx), _dp(theta),
           _dp(_f64(rad)) if rad is not None else nul, C.c_double(goal_weight), C.c_int(one_step_rule),
           _dp(x0a) if x0a is not None else nul, _dp(qa) if qa is not None else nul, _dp(sol),
           _dp(fres) if fres is not None else nul, _dp(obs) if obs is not None else nul,
           _dp(goal) if goal is not None else nul, _dp(cost) if cost is not None else nul, _ip(info), _ip(nfev),
           C.c_int(nthreads or len(os.sched_getaffinity(0))))
    if rc:
        raise NotImplementedError("unsupported obstacle in oracle")
    out = {"solution": sol, "info": info, "nfev": nfev}
    if want_residual:
        out["residual"] = fres
    if env is not None:
        out.update(obs=obs, goal=goal, cost=cost)
    return out


def curves_eval_batch(env, g, offsets, tube_num, rad, goal_weight, nthreads=None):
    """check_collision + SOAPWG own cost on GIVEN packed curves (g [nodes,4,4], offsets [B*T+1]) -> obs, goal, cost."""
    g = _f64(g).reshape(-1, 16)
    off = np.ascontiguousarray(offsets, np.int64)
    B = (off.size - 1) // tube_num
    obs, goal, cost = np.zeros(B), np.zeros(B), np.zeros(B)
    f = lib().orc_curves_eval_batch
    f.argtypes = None
    rc = f(C.c_void_p(env._h), C.c_int(B), C.c_int(tube_num), _dp(g), off.ctypes.data_as(C.POINTER(C.c_longlong)),
           _dp(_f64(rad)), C.c_double(goal_weight), _dp(obs), _dp(goal), _dp(cost),
           C.c_int(nthreads or len(os.sched_getaffinity(0))))
    if rc:
        raise NotImplementedError("unsupported obstacle in oracle")
    return obs, goal, cost


def static_solve_g(s, indices, thetas, initial_guess=None, xtol=0.0, one_step_rule=0):
    """Static.solve_g (static.py:71-135) -> (g list, solution x, nfev, info).  one_step_rule: see static_batch."""
    T = s.m.tube_num
    tot = sum(s.m.npts[n] for n in range(T))
    g = np.zeros((tot, 4, 4))
    counts = np.zeros(T, np.int32)
    sol = np.zeros(64)
    nfev, info = C.c_int(), C.c_int()
    x0 = None if initial_guess is None else _dp(_f64(initial_guess))
    lib().orc_static_solve_g_ex.argtypes = None
    lib().orc_static_solve_g_ex(C.byref(s), _ip(np.ascontiguousarray(indices, np.int32)), _dp(_f64(thetas)), x0,
                                C.c_double(xtol), C.c_int(one_step_rule), _dp(g), _ip(counts), _dp(sol), C.byref(nfev),
                                C.byref(info))
    n = lib().orc_static_ndof(C.byref(s), _ip(np.ascontiguousarray(indices, np.int32)), _dp(_f64(thetas)))
    return _split(g, counts.tolist(), (4, 4)), sol[:n].copy(), nfev.value, info.value
