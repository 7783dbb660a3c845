from .heuristic import FollowTheLeader, FollowTheLeaderTranslation, FollowTheLeaderWInsertion  # noqa: F401
