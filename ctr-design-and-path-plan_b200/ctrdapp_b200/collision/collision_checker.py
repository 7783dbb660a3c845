"""CollisionChecker over libctrb200: mirror of ctrdapp/collision/collision_checker.py.

    CollisionChecker(init_objects_file).check_collision(curve, rad) -> (obstacle_min, goal_dist)

plus the batch entry points the planner/optimiser call with stacked candidates:
    check_collision_batch(g, node_offsets, rad)            curves already on hand
    evaluate_batch(model, idx, theta, rad, ...)            fused g-curve + collision + cost
"""
import ctypes as C

import numpy as np

from .. import _lib as L
from .init_collision import add_goal, add_obstacles


class CollisionChecker:
    def __init__(self, init_objects_file, ctx=None):
        self.ctx = ctx or L.default_context()
        self.obstacles = add_obstacles(init_objects_file)  # collision_checker.py:31-33
        self.goal = add_goal(init_objects_file)
        self._h = C.c_void_p()
        self._build_env()

    @classmethod
    def from_objects(cls, obstacles, goal, ctx=None):
        self = cls.__new__(cls)
        self.ctx = ctx or L.default_context()
        self.obstacles = list(obstacles)
        self.goal = goal
        self._h = C.c_void_p()
        self._build_env()
        return self

    def _build_env(self):
        meshes = [o.tris.reshape(-1, 9) for o in self.obstacles if o.shape == L.SHAPE_MESH]
        tris = np.ascontiguousarray(np.concatenate(meshes, 0)) if meshes else np.zeros((0, 9))
        prims = [o.prim for o in self.obstacles if o.shape != L.SHAPE_MESH]
        arr = (L.Prim * max(1, len(prims)))(*prims)
        goal_prim, goal_tris = None, np.zeros((0, 9))
        if self.goal is not None:
            if self.goal.shape == L.SHAPE_MESH:
                goal_prim = L.Prim()
                goal_prim.shape = L.SHAPE_MESH
                goal_tris = np.ascontiguousarray(self.goal.tris.reshape(-1, 9))
            else:
                goal_prim = self.goal.prim
        L.check(L.lib().ctrb_env_create(self.ctx.handle, L.ptr(tris), len(tris), arr, len(prims),
                                        C.byref(goal_prim) if goal_prim is not None else None,
                                        L.ptr(goal_tris), len(goal_tris), C.byref(self._h)))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                L.lib().ctrb_env_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def stats(self):
        out = (C.c_int32 * 4)()
        L.check(L.lib().ctrb_env_stats(self._h, out))
        return {"bvh_nodes": out[0], "bvh_leaves": out[1], "bvh_depth": out[2], "triangles": out[3]}

    # ------------------------------------------------------------------ reference signature
    def check_collision(self, curve, rad):
        """collision_checker.py:35-70."""
        counts = [len(t) for t in curve]
        off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        parts = [np.asarray(t, np.float64).reshape(-1, 4, 4) for t in curve if len(t)]
        g = np.ascontiguousarray(np.concatenate(parts, 0)) if parts else np.zeros((0, 4, 4))
        obs, goal = self.check_collision_batch(g, off, rad, tube_num=len(curve))
        return float(obs[0]), float(goal[0])

    # ------------------------------------------------------------------ batch
    def check_collision_batch(self, g, node_offsets, rad, tube_num=None):
        tube_num = tube_num or len(rad)
        g = np.ascontiguousarray(g, np.float64)
        node_offsets = np.ascontiguousarray(node_offsets, np.int64)
        B = (len(node_offsets) - 1) // tube_num
        rad = np.ascontiguousarray(rad, np.float64)
        obs = np.empty(B)
        goal = np.empty(B)
        L.check(L.lib().ctrb_check_collision_batch(self.ctx.handle, self._h, B, tube_num, L.ptr(g),
                                                   L.ptr(node_offsets), L.ptr(rad), L.ptr(obs), L.ptr(goal),
                                                   L.MEM_HOST))
        return obs, goal

    def evaluate_batch(self, model, idx, theta, rad, heuristic_type="square_obstacle_avg_plus_weighted_goal",
                       goal_weight=1.0, only_tip=True, insertion_weight=0.0):
        """Fused hot path for B configurations (host numpy in/out):
        -> (obstacle_min [B], goal_dist [B], cost [B], collide [B] uint8)."""
        T = model.tube_num
        idx = np.ascontiguousarray(idx, np.int32).reshape(-1, T)
        theta = np.ascontiguousarray(theta, np.float64).reshape(-1, T)
        rad = np.ascontiguousarray(rad, np.float64)
        B = idx.shape[0]
        cd = L.CostDesc(L.COST_TYPES[heuristic_type], float(goal_weight), int(bool(only_tip)), float(insertion_weight))
        obs = np.empty(B)
        goal = np.empty(B)
        cost = np.empty(B)
        coll = np.empty(B, np.uint8)
        if getattr(model, "num_unknowns", None):  # static model: solve -> collide, plus the converged flag
            info = np.empty(B, np.int32)
            nfev = np.empty(B, np.int32)
            L.check(L.lib().ctrb_static_eval_batch(self.ctx.handle, model.handle, self._h, C.byref(cd), B, L.ptr(idx),
                                                   L.ptr(theta), L.ptr(rad), L.ptr(obs), L.ptr(goal), L.ptr(cost),
                                                   L.ptr(coll), L.ptr(info), L.ptr(nfev), L.MEM_HOST))
            self.last_static = {"info": info, "nfev": nfev, "converged": info == 1}
            return obs, goal, cost, coll
        L.check(L.lib().ctrb_eval_batch(self.ctx.handle, model.handle, self._h, C.byref(cd), B, L.ptr(idx),
                                        L.ptr(theta), L.ptr(rad), L.ptr(obs), L.ptr(goal), L.ptr(cost),
                                        L.ptr(coll), L.MEM_HOST))
        return obs, goal, cost, coll
