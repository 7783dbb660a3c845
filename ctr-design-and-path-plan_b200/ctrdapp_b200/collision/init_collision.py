"""Environment loading: mirror of ctrdapp/collision/init_collision.py.

The reference builds fcl.CollisionObjects; here parse_json returns plain
descriptions (triangle arrays / primitive records) that CollisionChecker
uploads into one ctrb_env (BVH built by the library).
"""
import json
import pathlib

import numpy as np

from .. import _lib as L


def parse_mesh(filename):
    """init_collision.py:59-88: (vertices, tris) of an STL file as numpy-stl's Mesh.points gives them:
    three un-deduplicated float32 vertices per triangle.  Binary STL is detected by size
    (84 + 50 n bytes), not by the leading word -- the colon files' binary header starts with 'solid'."""
    raw = pathlib.Path(filename).read_bytes()
    tri = None
    if len(raw) >= 84:
        n = int(np.frombuffer(raw, "<u4", 1, 80)[0])
        if 84 + 50 * n == len(raw):
            rec = np.frombuffer(raw, np.dtype([("n", "<f4", 3), ("v", "<f4", 9), ("a", "<u2")]), n, 84)
            tri = rec["v"].reshape(n, 3, 3)
    if tri is None:  # ASCII
        verts = [ln.split()[1:4] for ln in raw.decode("ascii", "replace").splitlines()
                 if ln.strip().startswith("vertex")]
        tri = np.asarray(verts, np.float32).reshape(-1, 3, 3)
    vertices = tri.reshape(-1, 3)
    tris = np.arange(len(vertices)).reshape(-1, 3).tolist()
    return vertices, tris


def rotate_align(vector_to, vector_from=np.array([0, 0, 1])):
    """init_collision.py:149-188 (Kevin Moran's no-acos alignment)."""
    vector_to = np.asarray(vector_to, dtype=float)
    vector_from = np.asarray(vector_from, dtype=float)
    axis = np.cross(vector_from, vector_to)
    cos_a = np.dot(vector_from, vector_to)
    k = 1 / (1 + cos_a)
    return np.array([[(axis[0] * axis[0] * k) + cos_a, (axis[1] * axis[0] * k) - axis[2], (axis[2] * axis[0] * k) + axis[1]],
                     [(axis[0] * axis[1] * k) + axis[2], (axis[1] * axis[1] * k) + cos_a, (axis[2] * axis[1] * k) - axis[0]],
                     [(axis[0] * axis[2] * k) - axis[1], (axis[1] * axis[2] * k) + axis[0], (axis[2] * axis[2] * k) + cos_a]])


class EnvObject:
    """What stands in for fcl.CollisionObject: either a translated triangle soup or a primitive."""

    def __init__(self, shape, tris=None, prim=None, description=None):
        self.shape = shape
        self.tris = tris      # [n,3,3] float64, translated
        self.prim = prim      # _lib.Prim
        self.description = description


def parse_json(object_type, init_objects_file):
    """init_collision.py:91-146."""
    init_objects_file = pathlib.Path(init_objects_file)
    with init_objects_file.open() as f:
        full_json = json.load(f)
    list_objects = full_json.get(object_type)

    collision_objects = []
    for o in list_objects:
        if o.get('mesh'):
            file = init_objects_file.parent / o.get('filename')
            vertices, _ = parse_mesh(str(file))
            # float32 vertex promoted to double, translation-only transform (init_collision.py:82-86,125)
            tris = vertices.astype(np.float64).reshape(-1, 3, 3) + np.asarray(o.get('transform'), np.float64)
            collision_objects.append(EnvObject(L.SHAPE_MESH, tris=tris, description=o))
            continue
        p = L.Prim()
        p.rot[:] = np.eye(3).ravel().tolist()
        if o.get('shape') == 'Sphere':
            p.shape = L.SHAPE_SPHERE
            p.dims[:] = [float(o.get('radius')), 0.0, 0.0]
        elif o.get('shape') == 'Cylinder':
            p.shape = L.SHAPE_CYLINDER
            p.dims[:] = [float(o.get('radius')), float(o.get('length')), 0.0]
            p.rot[:] = rotate_align(np.asarray(o.get('direction'))).ravel().tolist()
        elif o.get('shape') == 'Box':
            p.shape = L.SHAPE_BOX
            p.dims[:] = [float(o.get('x_length')), float(o.get('y_length')), float(o.get('z_length'))]
        else:
            raise NotImplementedError(f'Shape type {o.get("shape")} not supported.')
        p.center[:] = [float(x) for x in o.get('transform')]
        collision_objects.append(EnvObject(p.shape, prim=p, description=o))
    return collision_objects


def add_goal(init_objects_file):
    """init_collision.py:14-39: only the first goal entry is used."""
    goals = parse_json('goal', init_objects_file)
    return goals[0] if goals else None


def add_obstacles(init_objects_file):
    """init_collision.py:42-56."""
    return parse_json('obstacles', init_objects_file)
