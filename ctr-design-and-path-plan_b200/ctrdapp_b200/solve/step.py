"""Steering arithmetic of the planner (mirror of ctrdapp/solve/step.py), for one candidate or a stack of them.

A control point is [insertion_0, rotation_0, insertion_1, rotation_1, ...] (rrt.py:146-151).  The scalar
functions keep the reference's signatures; the `_batch` variants take [K, T] arrays and are what the batched
planner calls (same arithmetic, vectorised over candidates).
"""
from math import floor, pi

import numpy as np

TWO_PI = 2 * pi


def get_delta_rotation(prev, to):
    """Signed shortest rotation from prev to `to` on the circle [0, 2 pi] (step.py:103-125)."""
    gap = abs(to - prev)
    magnitude = min(gap, TWO_PI - gap)
    forward = (to > prev and (to - prev) < pi) or (prev > to and (prev - to) > pi)
    return magnitude if forward else -magnitude


def delta_rotation_batch(prev, to):
    prev, to = np.asarray(prev, np.float64), np.asarray(to, np.float64)
    gap = np.abs(to - prev)
    magnitude = np.minimum(gap, TWO_PI - gap)
    forward = ((to > prev) & ((to - prev) < pi)) | ((prev > to) & ((prev - to) > pi))
    return np.where(forward, magnitude, -magnitude)


def step(prev, control_params, max_bound):
    """Point at most max_bound (Euclidean, insertions only) from prev towards the control point (step.py:7-45)."""
    to = control_params[::2]
    if len(prev) != len(to):
        raise ValueError("Given vectors are not the same dimension")
    out = step_batch(np.asarray([prev], np.float64), np.asarray([to], np.float64), max_bound)[0]
    return out.tolist() if isinstance(to, list) else out


def step_batch(prev, to, max_bound):
    """prev, to: [K, T] insertions -> [K, T]."""
    vec = to - prev
    mag = np.sqrt(np.sum(vec * vec, axis=1))
    far = mag > max_bound
    safe = np.where(far, mag, 1.0)
    return np.where(far[:, None], vec * max_bound / safe[:, None] + prev, to)


def step_rotation(prev, control_params, rotation_max):
    """Next rotation (mod 2 pi) and the signed rotation taken, bounded by rotation_max in the Euclidean norm over
    tubes (step.py:48-100)."""
    to = control_params[1::2]
    if len(prev) != len(to):
        raise ValueError("Given vectors are not the same dimension")
    final, delta = step_rotation_batch(np.asarray([prev], np.float64), np.asarray([to], np.float64), rotation_max)
    return final[0].tolist(), delta[0].tolist()


def step_rotation_batch(prev, to, rotation_max):
    """prev, to: [K, T] rotations -> (new rotation [K, T], signed delta [K, T])."""
    vel = delta_rotation_batch(prev, to)
    mag = np.sqrt(np.sum(vel * vel, axis=1))
    far = mag > rotation_max
    safe = np.where(far, mag, 1.0)
    delta = np.where(far[:, None], vel * rotation_max / safe[:, None], vel)
    final = np.where(far[:, None], (delta + prev) % TWO_PI, to)
    return final, delta


def get_single_tube_value(tube_control_params, insert_neighbor, rotation_neighbor, neighbor_parent, probability,
                          random_num):
    """single_tube_control: keep the neighbour's values for every tube but one; the tube that moved last is
    chosen again with `probability`, any other with (1 - probability) / (T - 1) (step.py:128-189)."""
    T = len(insert_neighbor)
    if T == 1:
        return tube_control_params, 0
    moved = [k for k in range(T) if insert_neighbor[k] - neighbor_parent[k] != 0]
    if len(moved) > 1:
        raise ValueError("Multiple tubes inserted in neighbor node.")
    last = moved[0] if moved else 0
    others = [k for k in range(T) if k != last]
    shifted = random_num - probability
    chosen = last if shifted < 0 else others[floor(shifted / ((1 - probability) / (T - 1)))]
    out = []
    for k in range(T):
        if k == chosen:
            out += [tube_control_params[2 * k], tube_control_params[2 * k + 1]]
        else:
            out += [insert_neighbor[k], rotation_neighbor[k]]
    return out, chosen
