"""Heuristic factories (mirror of ctrdapp/heuristic/heuristic_factory.py): `create(**kwargs)` picks what each cost
needs out of (min_obstacle_distance, goal_distance, follow_the_leader, insertion_fraction), `create_from_old`
re-creates a node's cost under a new parent (rrt_star.py:86-87)."""
from .heuristic import (FollowTheLeader, FollowTheLeaderTranslation, FollowTheLeaderWInsertion, Heuristic,  # noqa: F401
                        OnlyGoalDistance, SquareObstacleAvgPlusWeightedGoal)


class HeuristicFactory:
    #: column of ctrb_eta_ftl_batch's ftl_mag this cost reads when only_tip / not only_tip (None: no FTL term)
    ftl_mag_columns = None
    needs_parent = False  # does the cost of a node depend on which parent it hangs from?

    def create(self, **kwargs) -> Heuristic:
        raise NotImplementedError

    def create_from_old(self, heuristic, **kwargs) -> Heuristic:
        raise NotImplementedError


class SOAPWGFactory(HeuristicFactory):
    def __init__(self, goal_weight, *args):
        self.goal_weight = goal_weight

    def create(self, **kwargs):
        return SquareObstacleAvgPlusWeightedGoal(self.goal_weight, kwargs.get('min_obstacle_distance'),
                                                 kwargs.get('goal_distance'))

    def create_from_old(self, heuristic, **kwargs):
        return self.create(min_obstacle_distance=heuristic.store_obstacle_min, goal_distance=heuristic.goal_dist)


class OnlyGoalDistanceFactory(HeuristicFactory):
    def __init__(self, *args):
        pass

    def create(self, **kwargs):
        return OnlyGoalDistance(kwargs.get('goal_distance'))

    def create_from_old(self, heuristic, **kwargs):
        return self.create(goal_distance=heuristic.get_cost())


class FTLFactory(HeuristicFactory):
    ftl_mag_columns = (0, 1)
    needs_parent = True
    _cls = FollowTheLeader

    def __init__(self, only_tip, *args):
        self.only_tip = only_tip

    def create(self, **kwargs):
        return self._cls(self.only_tip, kwargs.get('follow_the_leader'))

    def create_from_old(self, heuristic, **kwargs):
        return self.create(**kwargs)

    def create_from_magnitude(self, ftl_mag_row, insertion_fraction=None):
        """The same object from one row of ctrb_eta_ftl_batch's ftl_mag (no per-node arrays on the host)."""
        col = self.ftl_mag_columns[0 if self.only_tip else 1]
        return self._cls.from_magnitude(ftl_mag_row[col])


class FTLWInsertionFactory(FTLFactory):
    _cls = FollowTheLeaderWInsertion

    def __init__(self, only_tip, insertion_weight, *args):
        super().__init__(only_tip, *args)
        self.insertion_weight = insertion_weight

    def create(self, **kwargs):
        return FollowTheLeaderWInsertion(self.only_tip, self.insertion_weight, kwargs.get('follow_the_leader'),
                                         kwargs.get('insertion_fraction'))

    def create_from_old(self, heuristic, **kwargs):
        return self.create(follow_the_leader=kwargs.get('follow_the_leader'),
                           insertion_fraction=heuristic.insertion_fraction)

    def create_from_magnitude(self, ftl_mag_row, insertion_fraction=None):
        col = self.ftl_mag_columns[0 if self.only_tip else 1]
        return FollowTheLeaderWInsertion.from_magnitude(ftl_mag_row[col], self.insertion_weight, insertion_fraction)


class FTLTranslationFactory(FTLFactory):
    ftl_mag_columns = (2, 3)
    _cls = FollowTheLeaderTranslation


def create_heuristic_factory(configuration, heuristic_dict) -> HeuristicFactory:
    """heuristic_factory.py:173-203."""
    name = configuration.get("heuristic_type")
    params = [configuration.get(n) for n in heuristic_dict.get(name)]
    table = {"square_obstacle_avg_plus_weighted_goal": SOAPWGFactory, "only_goal_distance": OnlyGoalDistanceFactory,
             "follow_the_leader": FTLFactory, "follow_the_leader_w_insertion": FTLWInsertionFactory,
             "follow_the_leader_translation": FTLTranslationFactory}
    if name not in table:
        raise UserWarning(f"{name} is not a defined heuristic. Change config file.")
    return table[name](*params)
