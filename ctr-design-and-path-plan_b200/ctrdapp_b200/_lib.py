"""ctypes binding of include/ctrb200.h (the drop-in C ABI)."""
import ctypes as C
import pathlib
import threading

import numpy as np

MAX_TUBES = 4
MAX_DOF = 8
MEM_HOST, MEM_DEVICE = 0, 1
F64, F32 = 64, 32
DBL_MAX = float(np.finfo(np.float64).max)

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_NO_DEVICE, ERR_ALLOC = 0, -1, -2, -3, -4, -5

BASE_KINDS = {"constant": 0, "linear": 1, "quadratic": 2, "pure_helix": 3, "linear_helix": 4,
              "torsion_helix": 5, "torsion_linear_helix": 6, "full": 7}
SHAPE_NONE, SHAPE_SPHERE, SHAPE_BOX, SHAPE_CYLINDER, SHAPE_MESH = range(5)
COST_TYPES = {"square_obstacle_avg_plus_weighted_goal": 0, "only_goal_distance": 1, "follow_the_leader": 2,
              "follow_the_leader_w_insertion": 3, "follow_the_leader_translation": 4}


class CtrbError(RuntimeError):
    """Any non-zero status of libctrb200 that has no closer Python equivalent."""


class ModelDesc(C.Structure):
    _fields_ = [("model_type", C.c_int32), ("tube_num", C.c_int32),
                ("base_kind", C.c_int32 * MAX_TUBES), ("q_dof", C.c_int32 * MAX_TUBES),
                ("q", (C.c_double * MAX_DOF) * MAX_TUBES), ("tube_length", C.c_double * MAX_TUBES),
                ("delta_x", C.c_double), ("radius_outer", C.c_double * MAX_TUBES),
                ("radius_inner", C.c_double * MAX_TUBES), ("static_degree", C.c_int32),
                ("static_basis_type", C.c_int32)]


class Prim(C.Structure):
    _fields_ = [("shape", C.c_int32), ("center", C.c_double * 3), ("rot", C.c_double * 9),
                ("dims", C.c_double * 3)]


class CostDesc(C.Structure):
    _fields_ = [("cost_type", C.c_int32), ("goal_weight", C.c_double), ("only_tip", C.c_int32),
                ("insertion_weight", C.c_double)]


# every symbol include/ctrb200.h declares: name -> (restype, argtypes)
_vp = C.c_void_p
_i64 = C.c_int64
SYMBOLS = {
    "ctrb_last_error": (C.c_char_p, []),
    "ctrb_version": (C.c_char_p, []),
    "ctrb_device_count": (C.c_int, []),
    "ctrb_ctx_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "ctrb_ctx_destroy": (None, [_vp]),
    "ctrb_ctx_synchronize": (C.c_int, [_vp]),
    "ctrb_ctx_set_stream": (C.c_int, [_vp, _vp]),
    "ctrb_ctx_set_option": (C.c_int, [_vp, C.c_int, C.c_int]),
    "ctrb_ctx_get_option": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_int)]),
    "ctrb_measure_fma_peak": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_double)]),
    "ctrb_ctx_launch_count": (_i64, [_vp]),
    "ctrb_ctx_mark": (C.c_int, [_vp, C.c_int]),
    "ctrb_ctx_elapsed_ms": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "ctrb_ctx_last_kernel_ms": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_float)]),
    "ctrb_static_last_fallbacks": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "ctrb_static_last_fallback_reasons": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "ctrb_ctx_work_counters": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "ctrb_ctx_device_status": (C.c_int, [_vp, C.POINTER(C.c_int32)]),
    "ctrb_model_create": (C.c_int, [_vp, C.POINTER(ModelDesc), C.POINTER(_vp)]),
    "ctrb_model_destroy": (None, [_vp]),
    "ctrb_model_num_points": (C.c_int, [_vp, C.POINTER(C.c_int32)]),
    "ctrb_env_create": (C.c_int, [_vp, _vp, C.c_int32, C.POINTER(Prim), C.c_int32, C.POINTER(Prim), _vp, C.c_int32,
                                  C.POINTER(_vp)]),
    "ctrb_env_destroy": (None, [_vp]),
    "ctrb_env_stats": (C.c_int, [_vp, C.POINTER(C.c_int32)]),
    "ctrb_calculate_indices_batch": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, C.c_int]),
    "ctrb_g_node_offsets": (C.c_int, [_vp, _vp, _i64, _vp, C.c_int, _vp, C.c_int]),
    "ctrb_solve_g_batch": (C.c_int, [_vp, _vp, _i64, _vp, _vp, C.c_int, C.c_int, _vp, _vp, C.c_int]),
    "ctrb_eta_ftl_batch": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int]),
    "ctrb_static_num_unknowns": (C.c_int, [_vp, C.POINTER(C.c_int32)]),
    "ctrb_static_initial_guess": (C.c_int, [_vp, _vp]),
    "ctrb_static_residual_batch": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, C.c_int]),
    "ctrb_static_solve_g_batch": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp, C.c_int, C.c_double, _vp, _vp, _vp, _vp, _vp,
                                            C.c_int]),
    "ctrb_static_curves_batch": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, C.c_int]),
    "ctrb_static_eval_batch": (C.c_int, [_vp, _vp, _vp, C.POINTER(CostDesc), _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                         _vp, _vp, C.c_int]),
    "ctrb_check_collision_batch": (C.c_int, [_vp, _vp, _i64, C.c_int32, _vp, _vp, _vp, _vp, _vp, C.c_int]),
    "ctrb_eval_batch": (C.c_int, [_vp, _vp, _vp, C.POINTER(CostDesc), _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                  C.c_int]),
    "ctrb_knn_batch": (C.c_int, [_vp, C.c_int32, _i64, _vp, _i64, _vp, _vp, _vp, C.c_int32, _vp, _vp, C.c_int]),
    "ctrb_argmin_cost": (C.c_int, [_vp, _i64, _vp, _i64, C.POINTER(C.c_double), C.POINTER(_i64), C.c_int]),
    "ctrb_argmin_cost_device": (C.c_int, [_vp, _i64, _vp, _i64, _vp]),
    "ctrb_key_to_cost": (C.c_double, [_i64]),
    "ctrb_cost_to_key": (_i64, [C.c_double]),
}

_lib = None
_lock = threading.Lock()


def lib_path():
    """In-tree library; CTRB200_LIB overrides it (used by tools/ to compare build variants)."""
    import os
    if os.environ.get("CTRB200_LIB"):
        return pathlib.Path(os.environ["CTRB200_LIB"])
    return pathlib.Path(__file__).resolve().parent.parent / "libctrb200.so"


def lib():
    """Load libctrb200.so.  Raises if it has not been built: there is no fallback path."""
    global _lib
    with _lock:
        if _lib is None:
            p = lib_path()
            if not p.exists():
                raise CtrbError(f"{p} is missing: build it with `make -C {p.parent / 'csrc'}` "
                                "(python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback")
            h = C.CDLL(str(p))
            for name, (res, args) in SYMBOLS.items():
                fn = getattr(h, name)
                fn.restype = res
                fn.argtypes = args
            _lib = h
    return _lib


def check(rc):
    if rc == OK:
        return
    msg = lib().ctrb_last_error().decode("utf-8", "replace")
    if rc == ERR_INVALID:
        raise ValueError(msg)
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise CtrbError(f"libctrb200 status {rc}: {msg}")


def ptr(a):
    """void* of a numpy array (host) / torch tensor (device) / int / None."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError(f"cannot take a pointer of {type(a)}")


class Context:
    """ctrb_ctx: one device, one stream, scratch buffers."""

    def __init__(self, device=0):
        self._h = _vp()
        check(lib().ctrb_ctx_create(int(device), C.byref(self._h)))
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            lib().ctrb_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def synchronize(self):
        check(lib().ctrb_ctx_synchronize(self._h))

    def set_stream(self, cuda_stream):
        """Launch on a caller-owned cudaStream_t (int handle); None restores the ctx's own stream."""
        check(lib().ctrb_ctx_set_stream(self._h, C.c_void_p(cuda_stream) if cuda_stream else None))

    def measure_fma_peak(self, precision=64):
        out = C.c_double()
        check(lib().ctrb_measure_fma_peak(self._h, precision, C.byref(out)))
        return out.value

    def launch_count(self):
        return int(lib().ctrb_ctx_launch_count(self._h))

    def mark(self, which):
        check(lib().ctrb_ctx_mark(self._h, which))

    def elapsed_ms(self):
        ms = C.c_float()
        check(lib().ctrb_ctx_elapsed_ms(self._h, C.byref(ms)))
        return ms.value

    KERNELS = {"chain_eval": 0, "static_solve": 1, "gcurve": 2, "eta_ftl": 3, "static_fast": 4, "static_qr": 5,
               "static_curves": 6}
    # ctrb_option (include/ctrb200.h): run-time switches for parity tests and measurement
    OPTIONS = {"static_solver": 0, "static_one_step": 1, "fast_generic_n": 2, "fast_block_invert": 3, "eval_group": 4,
               "gcurve_blocks": 5, "max_nodes_hint": 6, "static_curves_scan": 7}

    def set_option(self, name, value):
        check(lib().ctrb_ctx_set_option(self._h, self.OPTIONS[name], int(value)))

    def get_option(self, name):
        v = C.c_int()
        check(lib().ctrb_ctx_get_option(self._h, self.OPTIONS[name], C.byref(v)))
        return v.value

    def options(self, **kw):
        """context manager: set options, restore the previous values on exit"""
        import contextlib

        @contextlib.contextmanager
        def _cm():
            old = {k: self.get_option(k) for k in kw}
            try:
                for k, v in kw.items():
                    self.set_option(k, v)
                yield self
            finally:
                for k, v in old.items():
                    self.set_option(k, v)
        return _cm()

    def last_kernel_ms(self, name):
        ms = C.c_float()
        check(lib().ctrb_ctx_last_kernel_ms(self._h, self.KERNELS[name], C.byref(ms)))
        return ms.value

    def static_fallbacks(self):
        n = C.c_uint64()
        check(lib().ctrb_static_last_fallbacks(self._h, C.byref(n)))
        return int(n.value)

    def device_status(self):
        f = C.c_int32()
        check(lib().ctrb_ctx_device_status(self._h, C.byref(f)))
        return int(f.value)

    def static_fallback_reasons(self):
        out = (C.c_uint64 * 6)()
        check(lib().ctrb_static_last_fallback_reasons(self._h, out))
        return dict(zip(("degenerate_section", "nonfinite_initial_residual", "singular_first_jacobian", "singular_later_jacobian",
                         "nonfinite_trial_point", "singular_broyden_update"), (int(v) for v in out)))

    def work_counters(self):
        out = (C.c_uint64 * 4)()
        check(lib().ctrb_ctx_work_counters(self._h, out))
        return {"bvh_nodes": out[0], "prefilter_tests": out[1], "exact_tests": out[2], "configs": out[3]}


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]
