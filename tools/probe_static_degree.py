"""Probe (dev container only: imports the reference from /root/reference): what does the reference's static model do for
`degree` != 2 under basis_type 'last_strain_base'?  Output committed as profiles/r2_probe_static_degree.txt."""
import contextlib
import io
import sys
import traceback

import numpy as np

sys.path.insert(0, "/root/reference")
from ctrdapp.model.model_factory import create_model  # noqa: E402

q = [0.02, 0.001, 0.05, 0.0001, 0.002]
for degree in (0, 1, 2, 3, 4):
    cfg = dict(model_type="static", tube_number=2, strain_bases=["linear", "linear"], q_dof=[2, 3], tube_lengths=[120, 120],
               delta_x=1, tube_radius={"outer": [0.9, 0.6], "inner": [0.7, 0.4]}, degree=degree, basis_type="last_strain_base")
    print(f"---- degree {degree}")
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            m = create_model(cfg, q)
        print("  unknowns per block:", m.ndof, " initial guess length", len(m.general_init_guess))
        b = m.static_basis.get_basis(2, 2, 2, {(2, 2): 0}, [0, 0])
        print("  tube==section basis at x = 2: shape", b.shape)
        for r in range(3):
            print("    row", r, np.array2string(b[r], precision=3))
        print("  all-zero columns:", [int(c) for c in np.flatnonzero(~b.any(axis=0))],
              " columns with two non-zeros:", [int(c) for c in np.flatnonzero((b != 0).sum(axis=0) > 1)])
        out = io.StringIO()
        with contextlib.redirect_stdout(out):
            g = m.solve_g(indices=[60, 30], thetas=[0.3, 1.2])
        tip = np.asarray(g[-1][-1])[0:3, 3]
        print("  solve_g([60, 30], [0.3, 1.2]) ->", [len(t) for t in g], "nodes, tip", np.array2string(tip, precision=4),
              "| first line printed by the reference (nfev):", out.getvalue().splitlines()[0] if out.getvalue() else "")
    except Exception as e:  # noqa: BLE001
        print("  FAILS:", type(e).__name__, str(e).splitlines()[0])
        tb = traceback.extract_tb(e.__traceback__)[-1]
        print("   at", tb.filename.replace("/root/reference/", ""), "line", tb.lineno, ":", tb.line)
