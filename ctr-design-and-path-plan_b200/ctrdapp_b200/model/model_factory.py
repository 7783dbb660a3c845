"""create_model(configuration, q): same contract as ctrdapp/model/model_factory.py:12-78 -- a flat q (or one array
per tube), scalar-or-list `q_dof` / `tube_lengths`, ValueError for a q of the wrong size, UserWarning for an unknown
model type -- returning a model backed by a libctrb200 handle."""
import numpy as np

from .kinematic import Kinematic
from .model import Model


def _per_tube(value, tube_num, what):
    """A scalar configuration entry applies to every tube (model_factory.py:33-38, 58-63)."""
    if isinstance(value, int):
        print(f'{what}: {value}.')
        return [value] * tube_num
    return value


def _split_q(q, q_dof, tube_num):
    if len(q) == sum(q_dof) and not isinstance(q[0], list):      # one flat list over all tubes
        bounds = np.concatenate([[0], np.cumsum(q_dof)])
        return [np.asarray(q[bounds[n]:bounds[n + 1]]) for n in range(tube_num)]
    if len(q) == tube_num:                                       # already one entry per tube
        return [np.asarray(tube_q) for tube_q in q]
    raise ValueError(f"Given q of {q} is not the correct size.")


def create_model(configuration, q, ctx=None) -> Model:
    kind = configuration.get("model_type")
    tube_num = configuration.get('tube_number')
    q_dof = _per_tube(configuration.get('q_dof'), tube_num, 'Each base has same degrees of freedom')
    lengths = _per_tube(configuration.get('tube_lengths'), tube_num, 'Each tube length is the same')
    args = (tube_num, _split_q(q, q_dof, tube_num), q_dof, lengths, configuration.get('delta_x'),
            configuration.get('strain_bases'))
    if kind == "kinematic":
        return Kinematic(*args, ctx=ctx)
    if kind == "static":
        from .static import Static
        return Static(*args, configuration.get('tube_radius'), configuration.get('degree'),
                      configuration.get('basis_type'), ctx=ctx)
    raise UserWarning(f"{kind} is not a defined model. Change config file.")
