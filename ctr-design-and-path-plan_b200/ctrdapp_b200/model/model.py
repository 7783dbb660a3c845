"""Abstract model: mirrors ctrdapp/model/model.py:4-102 (same attributes and method names)."""
from abc import ABC, abstractmethod


class Model(ABC):
    def __init__(self, tube_num, max_tube_length, delta_x):
        self.tube_num = tube_num
        self.max_tube_length = max_tube_length
        self.delta_x = delta_x
        # model.py:37-38 (Python round = half to even; the library uses nearbyint)
        self.num_discrete_points = [round(length / delta_x) + 1 for length in max_tube_length]

    @abstractmethod
    def solve_integrate(self, delta_theta, delta_insertion, this_theta, this_insertion, prev_g, invert_insert=True,
                        need_g_out=True):
        pass

    @abstractmethod
    def solve_g(self, indices=None, thetas=None, **kwargs):
        pass
