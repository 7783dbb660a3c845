"""Strain-base names and their degree-of-freedom rules (ctrdapp/model/strain_bases.py:227-351).
The bases themselves are evaluated on the device (csrc/ctrb_se3.cuh: strain_omega)."""
from .._lib import BASE_KINDS


def check_qdof(base, q_dof):
    """strain_bases.py:316-351: same messages, same ValueError."""
    if base in ("linear_helix", "pure_helix"):
        if q_dof != 2:
            raise ValueError(f'{base} should have 2 degrees of freedom, not {q_dof}.')
    elif base in ("quadratic", "linear"):
        if q_dof < 2 or q_dof > 3:
            raise ValueError(f'{base} should have 2 or 3 degrees of freedom, not {q_dof}.')
    elif base == "constant":
        if q_dof != 1:
            raise ValueError(f'{base} should have 1 degrees of freedom, not {q_dof}.')
    elif base in ("torsion_helix", "torsion_linear_helix"):
        if q_dof != 3:
            raise ValueError(f'{base} should have 3 degrees of freedom, not {q_dof}.')
    elif base == "full":
        if q_dof < 5 or q_dof > 8:
            raise ValueError(f'{base} should have 5-8 degrees of freedom, not {q_dof}.')
    else:
        print(f'{base} is not a defined strain base.')


def get_strains(names, q_dof):
    """strain_bases.py:227-268; returns the validated base names (the device holds the functions)."""
    out = []
    for n, d in zip(names, q_dof):
        check_qdof(n, d)
        if n in BASE_KINDS:
            out.append(n)
        else:
            print(f'{n} is not a defined strain base.')
    return out


def max_from_base(base, q_max, length, q_dof):
    """strain_bases.py:271-313 (used by the optimiser's initial simplex)."""
    check_qdof(base, q_dof)
    tube_array = []
    if base in ("linear", "quadratic"):
        reduced_max = q_max / length if base == "linear" else q_max / length ** 2
        if q_dof >= 3:
            tube_array = [q_max, -reduced_max]
        elif q_dof == 2:
            tube_array = [q_max]
        tube_array.append(reduced_max)
    elif base == "linear_helix":
        tube_array = [q_max, q_max / length]
    elif base == "torsion_linear_helix":
        tube_array = [q_max, q_max, q_max / length]
    else:
        tube_array = [q_max] * q_dof
    return tube_array
