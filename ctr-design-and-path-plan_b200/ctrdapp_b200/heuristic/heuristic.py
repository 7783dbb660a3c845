"""Per-node cost objects of the planner (mirror of ctrdapp/heuristic/*.py; host-side bookkeeping only).

Every cost of the reference is "own term of this node, folded into a running statistic along the path from the
root": a running MEAN by generation for the obstacle term (obstacles_and_goal.py:47-88), a running SUM for the
follow-the-leader magnitudes (follow_the_leader.py:44-75), nothing for the goal distance.  The fold lives once in
`PathCost`; the concrete classes keep the reference's names and attribute names because the solver, the tests and
the result writers read them (`avg_obstacle_min`, `ftl_sum`, `generation`, ...).

The own terms themselves (1/d^2, screw magnitudes) are also what the fused device kernels emit
(`ctrb_eval_batch` cost, `ctrb_eta_ftl_batch` ftl_mag), so a batched planner can build these objects from
kernel outputs without touching the per-node arrays: see `from_magnitude`.
"""
import math
from abc import ABC, abstractmethod

import numpy as np


class Heuristic(ABC):
    """Interface of heuristic.py:4-60."""

    @abstractmethod
    def calculate_cost_from_parent(self, parent, reset=False, init_insertion=False):
        ...

    @abstractmethod
    def test_cost_from_parent(self, parent):
        ...

    @abstractmethod
    def get_cost(self):
        ...


class PathCost(Heuristic):
    """own value + (parent statistic, parent generation) -> statistic of this node."""

    generation = 0

    # -- to be provided by subclasses
    def _own(self):
        raise NotImplementedError

    def _get_stat(self):
        raise NotImplementedError

    def _set_stat(self, v):
        raise NotImplementedError

    @staticmethod
    def _fold(parent_stat, parent_gen, own):
        raise NotImplementedError

    def _restart(self):
        """all tubes back at zero insertion: the path statistic starts over (dynamic_tree.py:144-145)"""
        self.generation = 0

    def calculate_cost_from_parent(self, parent, reset=False, init_insertion=False):
        if self.generation != 0 and not reset:
            print("Cost already calculated. Do not run method twice.")
            return
        self._set_stat(self._fold(parent._get_stat(), parent.generation, self._own()))
        self.generation = parent.generation + 1
        if init_insertion:
            self._restart()

    def test_cost_from_parent(self, parent):
        keep = self._get_stat()
        self._set_stat(self._fold(parent._get_stat(), parent.generation, self._own()))
        cost = self.get_cost()
        self._set_stat(keep)
        return cost


class SquareObstacleAvgPlusWeightedGoal(PathCost):
    """goal_weight * goal_dist + mean over the path of (1 / obstacle_min)^2   (obstacles_and_goal.py:39-88)."""

    def __init__(self, goal_weight, this_obstacle_min, goal_dist):
        self.store_obstacle_min = this_obstacle_min
        self.this_obstacle_min = (1 / this_obstacle_min) ** 2
        self.avg_obstacle_min = self.this_obstacle_min
        self.generation = 0
        self.goal_weight = goal_weight
        self.goal_dist = goal_dist

    def _own(self):
        return self.this_obstacle_min

    def _get_stat(self):
        return self.avg_obstacle_min

    def _set_stat(self, v):
        self.avg_obstacle_min = v

    @staticmethod
    def _fold(parent_avg, parent_gen, own):
        return (parent_gen * parent_avg + own) / (parent_gen + 1)

    _calculate_avg = _fold  # the reference's name (obstacles_and_goal.py:67)

    def get_cost(self):
        return self.goal_dist * self.goal_weight + self.avg_obstacle_min


class OnlyGoalDistance(Heuristic):
    """cost = goal distance, independent of the path (only_goal_distance.py:20-40)."""

    def __init__(self, goal_dist):
        self.cost = goal_dist

    def get_cost(self):
        return self.cost

    def calculate_cost_from_parent(self, parent, reset=False, init_insertion=False):
        pass

    def test_cost_from_parent(self, parent):
        print("Parent doesn't effect cost of Only Goal Distance heuristic.")
        return self.cost


class FollowTheLeader(PathCost):
    """Running sum of follow-the-leader screw magnitudes (follow_the_leader.py:29-116)."""

    def __init__(self, only_tip, follow_the_leader):
        if only_tip:
            mag = self.calculate_magnitude(follow_the_leader[-1][-1])
        else:
            mag = np.mean([self.calculate_magnitude(a) for tube in follow_the_leader for a in tube])
        self._init_from_magnitude(mag)

    def _init_from_magnitude(self, mag):
        self.ftl_magnitude = mag
        self.ftl_sum = mag
        self.generation = 0

    @classmethod
    def from_magnitude(cls, magnitude, *args):
        """Build from a magnitude the device already reduced (ctrb_eta_ftl_batch ftl_mag columns: [0] tip /
        [1] mean screw magnitude, [2] tip / [3] mean translation magnitude); *args are the extra constructor
        arguments of the subclass after (only_tip, ..., follow_the_leader)."""
        self = cls.__new__(cls)
        self._init_extra(*args)
        self._init_from_magnitude(float(magnitude))
        return self

    def _init_extra(self):
        pass

    def _own(self):
        return self.ftl_magnitude

    def _get_stat(self):
        return self.ftl_sum

    def _set_stat(self, v):
        self.ftl_sum = v

    @staticmethod
    def _fold(parent_sum, parent_gen, own):
        return parent_sum + own

    _calculate_avg = _fold  # the reference's name (follow_the_leader.py:63); it sums

    def _restart(self):
        self.generation = 0
        self.ftl_sum = self.ftl_magnitude

    def get_cost(self):
        return self.ftl_sum

    def get_own_cost(self):
        return self.ftl_magnitude

    @staticmethod
    def calculate_magnitude(ftl_array):
        """|omega| of the screw [omega; v], or |v| for a pure translation (follow_the_leader.py:89-116)."""
        w = math.sqrt(float(ftl_array[0]) ** 2 + float(ftl_array[1]) ** 2 + float(ftl_array[2]) ** 2)
        if w == 0:
            return math.sqrt(float(ftl_array[3]) ** 2 + float(ftl_array[4]) ** 2 + float(ftl_array[5]) ** 2)
        return w


class FollowTheLeaderWInsertion(FollowTheLeader):
    """ftl_sum + insertion_weight * insertion_fraction (follow_the_leader.py:119-170)."""

    def __init__(self, only_tip, insertion_weight, follow_the_leader, insertion_fraction):
        self._init_extra(insertion_weight, insertion_fraction)
        super().__init__(only_tip, follow_the_leader)

    def _init_extra(self, insertion_weight, insertion_fraction):
        self.insertion_weight = insertion_weight
        self.insertion_fraction = insertion_fraction

    def get_cost(self):
        return self.ftl_sum + self.insertion_weight * self.insertion_fraction


class FollowTheLeaderTranslation(FollowTheLeader):
    """Magnitude = |v| only (follow_the_leader.py:173-189)."""

    @staticmethod
    def calculate_magnitude(ftl_array):
        return math.sqrt(float(ftl_array[3]) ** 2 + float(ftl_array[4]) ** 2 + float(ftl_array[5]) ** 2)
