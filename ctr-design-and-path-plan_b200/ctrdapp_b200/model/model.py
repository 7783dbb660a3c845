"""What every model handle exposes to the planner (the attribute and method names of ctrdapp/model/model.py:4-102,
because rrt.py, neldermead.py and the scripts read them): geometry of the discretisation on the host, everything
else behind a libctrb200 handle in the subclasses."""


def discrete_points(length, delta_x):
    """Nodes of one tube: round(L / dx) + 1 with Python's round-half-to-even (model.py:37-38); the library's
    ctrb_model_num_points applies nearbyint and the constructor of Kinematic asserts that both agree."""
    return round(length / delta_x) + 1


class Model:
    tube_num = 0
    delta_x = 1.0

    def __init__(self, tube_num, max_tube_length, delta_x):
        self.tube_num, self.max_tube_length, self.delta_x = tube_num, max_tube_length, delta_x
        self.num_discrete_points = [discrete_points(L, delta_x) for L in max_tube_length]

    # the two calls the planner makes; concrete models implement both (plus the *_batch variants)
    def solve_integrate(self, delta_theta, delta_insertion, this_theta, this_insertion, prev_g, invert_insert=True,
                        need_g_out=True):
        raise NotImplementedError(f"{type(self).__name__}.solve_integrate")

    def solve_g(self, indices=None, thetas=None, **kwargs):
        raise NotImplementedError(f"{type(self).__name__}.solve_g")
