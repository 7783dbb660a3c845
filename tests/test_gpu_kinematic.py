"""GPU parity: kinematic model kernels (through the C ABI) vs the reference's golden vectors
and vs the CPU oracle on seeded random batches."""
import math

import numpy as np
import pytest

from models_def import MODELS
from util import gpu_model, orc_model, rel_err, random_configs

pytestmark = pytest.mark.gpu


def _split(flat, counts):
    out, o = [], 0
    for c in counts:
        out.append(flat[o:o + c])
        o += c
    return out


@pytest.mark.parametrize("name", list(MODELS))
def test_solve_g_golden(golden_kin, name):
    """fp64 g-curves within 1e-10 relative of Kinematic.solve_g (north-star tolerance)."""
    m = gpu_model(name)
    assert m.num_discrete_points == golden_kin[f"{name}/npts"].tolist()
    flat = golden_kin[f"{name}/g_flat"]
    o = 0
    for idx, th, full, cnt in zip(golden_kin[f"{name}/g_idx"], golden_kin[f"{name}/g_theta"],
                                  golden_kin[f"{name}/g_full"], golden_kin[f"{name}/g_counts"]):
        g = m.solve_g(idx.tolist(), th.tolist(), full=bool(full))
        assert [len(t) for t in g] == cnt.tolist()
        ref = flat[o:o + cnt.sum()]
        o += cnt.sum()
        got = np.concatenate([np.asarray(t) for t in g], 0)
        assert rel_err(got, ref) < 1e-10
        assert np.all(got[:, 3, :] == [0, 0, 0, 1])


@pytest.mark.parametrize("name", list(MODELS))
def test_solve_integrate_golden(golden_kin, name):
    m = gpu_model(name)
    for c in range(4):
        k = f"{name}/si{c}"
        prev_g = _split(golden_kin[k + "/prev_g"], golden_kin[k + "/prev_counts"])
        d_th, d_ins, th, ins = golden_kin[k + "/args"]
        g, eta, nidx, tins, ftl = m.solve_integrate(list(d_th), list(d_ins), list(th), list(ins), prev_g)
        assert nidx == golden_kin[k + "/new_idx"].tolist()
        np.testing.assert_allclose(tins, golden_kin[k + "/true_ins"], rtol=0, atol=1e-12)
        assert [len(t) for t in g] == golden_kin[k + "/g_counts"].tolist()
        assert rel_err(np.concatenate([np.asarray(t) for t in g], 0), golden_kin[k + "/g"]) < 1e-10
        np.testing.assert_allclose(np.array([e[0] for e in eta]), golden_kin[k + "/eta"], rtol=1e-10, atol=1e-10)
        assert [len(t) for t in ftl] == golden_kin[k + "/ftl_counts"].tolist()
        np.testing.assert_allclose(np.concatenate([np.asarray(t) for t in ftl], 0), golden_kin[k + "/ftl"],
                                   rtol=1e-10, atol=1e-10)


def test_ftl_magnitudes_golden(golden_kin):
    """follow_the_leader.py:29-38,89-116,173-189 epilogue of the eta kernel."""
    name = "c3_three_tube"
    m = gpu_model(name)
    for c in range(4):
        k = f"{name}/si{c}"
        d_th, d_ins, th, ins = golden_kin[k + "/args"]
        T = m.tube_num
        this_ins = [L - i for i, L in zip(ins, m.max_tube_length)]
        prev_ins = [t + d for t, d in zip(this_ins, d_ins)]
        p, q = m.calculate_indices_batch([prev_ins], [this_ins])
        off = np.concatenate([[0], np.cumsum(golden_kin[k + "/prev_counts"])])
        _, _, mag = m.eta_ftl_batch(p, q, [d_th], golden_kin[k + "/prev_g"], off, want_ftl=False)
        np.testing.assert_allclose(mag[0], golden_kin[k + "/ftl_mag"], rtol=1e-9, atol=1e-12)


def test_calculate_indices_golden(golden_kin):
    from ctrdapp_b200.model.kinematic import calculate_indices
    rows = golden_kin["calc_idx/in"]
    want = golden_kin["calc_idx/out"]
    # batch per delta_x value (one model handle each)
    for dx in np.unique(rows[:, 6]):
        sel = rows[:, 6] == dx
        from ctrdapp_b200.model.kinematic import Kinematic
        m = Kinematic(2, [np.zeros(1)] * 2, [1, 1], [50.0, 80.0], float(dx), ["constant"] * 2)
        p, q = m.calculate_indices_batch(rows[sel, 0:2], rows[sel, 2:4])
        assert np.array_equal(np.concatenate([p, q], 1), want[sel])
    for args in [(0, 0.49, 0, 1), (0.49, 0, 1, 0), (-5, 0.49, 0, 1), (49.51, 50, 49, 50), (60, 49.05, 50, 49),
                 (10.3, 10.2, 11, 10), (19.49, 21.51, 19, 21), (21.51, 19.49, 22, 20), (10, 10, 10, 10)]:
        assert calculate_indices([args[0]], [args[1]], [50], 1.0) == ([args[2]], [args[3]])
    assert calculate_indices([10.27], [10.31], [50], 0.1) == ([103], [104])


@pytest.mark.parametrize("name", ["config_yaml", "c3_three_tube", "mixed_bases", "torsion_constant"])
def test_solve_g_batch_vs_oracle(name):
    from oracle import oracle as orc
    m, om = gpu_model(name), orc_model(name)
    rng = np.random.default_rng(1002)
    B = 600
    idx, theta = random_configs(m.num_discrete_points, B, rng)
    idx[0] = 0
    idx[1] = np.asarray(m.num_discrete_points) - 1  # nothing exposed: one node per tube
    for full in (False, True):
        g, off = m.solve_g_batch(idx, theta, full=full)
        g32, _ = m.solve_g_batch(idx, theta, full=full, precision=32)
        worst, worst32 = 0.0, 0.0
        for b in range(0, B, 7):
            ref = np.concatenate(orc.solve_g(om, idx[b].tolist(), theta[b].tolist(), full=full), 0)
            lo, hi = off[b * m.tube_num], off[(b + 1) * m.tube_num]
            assert hi - lo == len(ref)
            worst = max(worst, rel_err(g[lo:hi], ref))
            worst32 = max(worst32, float(np.max(np.linalg.norm(g32[lo:hi, 0:3, 3].astype(np.float64) - ref[:, 0:3, 3], axis=1))))
        assert worst < 1e-10, worst
        assert worst32 < 1e-4, worst32  # fp32 path: tip (every node) within 1e-4 mm


def test_eta_ftl_batch_vs_oracle():
    from oracle import oracle as orc
    name = "c3_three_tube"
    m, om = gpu_model(name), orc_model(name)
    rng = np.random.default_rng(9)
    B = 200
    N = np.asarray(m.num_discrete_points)
    prev_idx = np.stack([rng.integers(2, N[t] - 2, B) for t in range(3)], 1).astype(np.int32)
    new_idx = (prev_idx + rng.integers(-2, 3, (B, 3))).astype(np.int32)
    theta = rng.uniform(0, 2 * np.pi, (B, 3))
    dth = rng.uniform(-0.17, 0.17, (B, 3))
    pg, poff = m.solve_g_batch(prev_idx, theta, full=False)
    eta, ftl, mag = m.eta_ftl_batch(prev_idx, new_idx, dth, pg, poff)
    for b in range(0, B, 5):
        prev = [pg[poff[b * 3 + n]:poff[b * 3 + n + 1]] for n in range(3)]
        vel = (prev_idx[b] - new_idx[b]) * m.delta_x
        e_ref, f_ref = orc.solve_eta(om, vel, prev_idx[b], dth[b], prev)
        np.testing.assert_allclose(eta[b], np.array([e[0] for e in e_ref]), rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(ftl[poff[b * 3]:poff[b * 3 + 3]], np.concatenate(f_ref, 0), rtol=1e-10, atol=1e-10)


def test_reference_unit_test_literals():
    """test/test_model.py:117-204 through the mirrored Kinematic class."""
    from ctrdapp_b200.model.kinematic import Kinematic
    no1 = Kinematic(1, [np.array([0, 0])], [2], [1], 0.1, ["linear"])
    fine = Kinematic(1, [np.array([0, 0])], [2], [1], 0.05, ["linear"])
    g, gd = no1.solve_g(full=True)[0], fine.solve_g(full=True)[0]
    end = np.eye(4)
    end[0, 3] = -1
    np.testing.assert_equal(g[-1], np.eye(4))
    np.testing.assert_almost_equal(g[0], end)
    for i in range(11):
        np.testing.assert_almost_equal(g[i], gd[i * 2])
    g = no1.solve_g(indices=[5], full=True)[0]
    np.testing.assert_equal(g[5], np.eye(4))
    g2pi = no1.solve_g(indices=[0], thetas=[2 * math.pi], full=True)[0]
    np.testing.assert_almost_equal(g2pi[0], np.eye(4))
    c1 = Kinematic(1, [np.array([1, 0])], [2], [math.pi / 2], math.pi / 20, ["linear"])
    g = c1.solve_g(indices=[0], full=True)[0]
    tip = np.array([[0, 0, 1, 1], [0, 1, 0, 0], [-1, 0, 0, -1], [0, 0, 0, 1.0]])
    np.testing.assert_equal(g[0], np.eye(4))
    np.testing.assert_almost_equal(g[-1], tip)
    c2 = Kinematic(2, [np.array([1, 0]), np.array([1, 0])], [2, 2], [math.pi / 2] * 2, math.pi / 20, ["linear"] * 2)
    g = c2.solve_g(indices=[0, 0], thetas=[0, math.pi / 2], full=True)
    np.testing.assert_almost_equal(g[1][-1], np.array([[0, 1, 0, 1], [1, 0, 0, 1], [0, 0, -1, -2], [0, 0, 0, 1.0]]))
    # FTL = 0 for constant curvature (test_integration.py:79-90)
    from oracle import oracle as orc
    mc = Kinematic(1, [np.array([0.05])], [1], [60], 0.5, ["constant"])
    prev = mc.solve_g(indices=[64], thetas=[-85], full=False)
    *_, ftl = mc.solve_integrate([0.85], [2], [0], [30], prev)
    mags = [np.linalg.norm(f[0:3]) if np.linalg.norm(f[0:3]) != 0 else np.linalg.norm(f[3:6]) for f in ftl[0]]
    assert abs(np.mean(mags) - 0.85) < 1e-7


def test_bad_arguments_raise():
    from ctrdapp_b200.model.kinematic import Kinematic
    with pytest.raises(ValueError):
        Kinematic(1, [np.array([0.1, 0.2])], [2], [10], 1, ["constant"])  # check_qdof (strain_bases.py:316-351)
    m = gpu_model("config_yaml")
    with pytest.raises(ValueError):
        m.solve_g_batch([[121, 0]], [[0.0, 0.0]])  # index beyond num_discrete_points - 1
    g, off = m.solve_g_batch(np.zeros((0, 2), np.int32), np.zeros((0, 2)))
    assert g.shape == (0, 4, 4) and off.tolist() == [0]


def test_eta_ftl_device_buffers_report_short_parents_through_the_status_flag():
    """ctrb_eta_ftl_batch with device buffers does not synchronise, so "parent curve shorter than npts - prev_idx" (an
    IndexError in the reference, CTRB_ERR_INVALID for host buffers) is reported through ctrb_ctx_device_status; the rows the
    kernel does not reach stay untouched (the Python mirror zero-fills them)."""
    import ctypes as C
    import torch
    from ctrdapp_b200 import _lib as L
    m = gpu_model("config_yaml")
    dev = torch.device("cuda", 0)
    B = 64
    prev_idx = np.full((B, 2), 100, np.int32)
    g, off = m.solve_g_batch(prev_idx, np.zeros((B, 2)), full=False)          # 21 nodes per tube
    new_idx = prev_idx - 1
    dth = np.full((B, 2), 0.05)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)          # noqa: E731
    gd, offd, pid, nid, dthd = t(g), t(off), t(prev_idx), t(new_idx), t(dth)
    eta = torch.zeros((B, 2, 6), dtype=torch.float64, device=dev)
    mag = torch.zeros((B, 4), dtype=torch.float64, device=dev)

    def call(pi):
        L.check(L.lib().ctrb_eta_ftl_batch(m.ctx.handle, m.handle, B, L.ptr(pi), L.ptr(nid), None, L.ptr(dthd), L.ptr(gd),
                                           L.ptr(offd), L.ptr(eta), None, L.ptr(mag), L.MEM_DEVICE))
        return m.ctx.device_status()
    assert call(pid) == 0
    ref = eta.clone()
    host_eta, _, host_mag = m.eta_ftl_batch(prev_idx, new_idx, dth, g, off, want_ftl=False)
    assert np.array_equal(ref.cpu().numpy(), host_eta) and np.array_equal(mag.cpu().numpy(), host_mag)
    bad = prev_idx.copy()
    bad[7, 1] = 90                                                            # needs 31 nodes, the parent curve has 21
    assert call(t(bad)) == 1
    with pytest.raises(ValueError, match="fewer nodes"):
        m.eta_ftl_batch(bad, new_idx, dth, g, off, want_ftl=False)           # host buffers: the call itself fails
    assert call(pid) == 0                                                     # the flag is per call
