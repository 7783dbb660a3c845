"""create_model: mirror of ctrdapp/model/model_factory.py:12-78."""
import numpy as np

from .kinematic import Kinematic
from .model import Model


def create_model(configuration, q, ctx=None) -> Model:
    name = configuration.get("model_type")
    tube_num = configuration.get('tube_number')
    names = configuration.get('strain_bases')

    config_dof = configuration.get('q_dof')
    if isinstance(config_dof, int):
        output_q_dof = [config_dof] * tube_num
        print(f'Each base has same degrees of freedom: {config_dof}.')
    else:
        output_q_dof = config_dof

    if len(q) == sum(output_q_dof) and not isinstance(q[0], list):  # single flat list
        q_list, dof_sum = [], 0
        for n in range(tube_num):
            q_list.append(np.asarray(q[dof_sum:dof_sum + output_q_dof[n]]))
            dof_sum += output_q_dof[n]
        output_q = q_list
    elif len(q) == tube_num:
        output_q = [np.asarray(this_q) for this_q in q]
    else:
        raise ValueError(f"Given q of {q} is not the correct size.")

    config_lengths = configuration.get('tube_lengths')
    if isinstance(config_lengths, int):
        lengths = [config_lengths] * tube_num
        print(f'Each tube length is the same: {config_lengths}.')
    else:
        lengths = config_lengths
    delta_x = configuration.get('delta_x')

    if name == "kinematic":
        return Kinematic(tube_num, output_q, output_q_dof, lengths, delta_x, names, ctx=ctx)
    elif name == "static":
        from .static import Static  # built in a later milestone
        return Static(tube_num, output_q, output_q_dof, lengths, delta_x, names, configuration.get('tube_radius'),
                      configuration.get('degree'), configuration.get('basis_type'), ctx=ctx)
    else:
        raise UserWarning(f"{name} is not a defined model. Change config file.")
