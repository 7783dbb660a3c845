"""CPU, world_size 2 over gloo: the multi-GPU host logic (sharding + best-node gather + cost gather)."""
import json
import os
import pathlib
import socket
import subprocess
import sys

import numpy as np
import pytest
from ctrdapp_b200 import dist as cdist


def test_shard_range_covers_batch():
    for total in (0, 1, 7, 10_000_001):
        for world in (1, 2, 3, 8):
            edges = [cdist.shard_range(total, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1


def test_local_best_ignores_discarded():
    assert cdist.local_best([np.nan, 3.0, 1.5, np.nan, 1.5], 100) == (1.5, 102)
    assert cdist.local_best([np.nan, np.nan]) == (float("inf"), -1)


def test_gloo_two_ranks_best_node(tmp_path):
    rng = np.random.default_rng(0)
    costs = rng.uniform(1, 100, 4096)
    costs[rng.random(4096) < 0.6] = np.nan   # colliding configurations are discarded (rrt.py:116)
    costs[3000] = 0.25
    costs[100] = 0.25                         # tie: the lowest index wins
    np.save(tmp_path / "costs.npy", costs)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    worker = pathlib.Path(__file__).with_name("_dist_worker.py")
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   GLOO_SOCKET_IFNAME="lo")
        procs.append(subprocess.Popen([sys.executable, str(worker), str(tmp_path / "costs.npy")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=240) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se[-2000:]
    want = cdist.local_best(costs)
    assert want == (0.25, 100)
    for so, _ in outs:
        res = json.loads(so.strip().splitlines()[-1])
        assert tuple(res["best"]) == want and res["costs_ok"]


def test_cost_key_is_order_preserving():
    """The packed key of the best-node reduction (ctrb_cost_to_key / dist.cost_to_key): signed 64-bit, a < b iff key(a) < key(b),
    exact round trip, NaN / inf (colliding, no candidate) above every finite cost; the C and the NumPy version agree."""
    from ctrdapp_b200 import _lib
    rng = np.random.default_rng(3)
    v = np.concatenate([rng.standard_normal(2000) * 10.0 ** rng.integers(-300, 300, 2000), [0.0, 1e-320, -1e-320, 1.0, -1.0]])
    k = cdist.cost_to_key(v)
    order = np.argsort(v, kind="stable")
    assert np.all(np.diff(k[order]) >= 0) and np.all((np.diff(k[order]) > 0) == (np.diff(v[order]) > 0))
    lib = _lib.lib()
    for x, kk in zip(v[:200], k[:200]):
        assert lib.ctrb_cost_to_key(float(x)) == int(kk)
        assert cdist.key_to_cost(int(kk)) == x and lib.ctrb_key_to_cost(int(kk)) == x
    assert cdist.cost_to_key(float("nan")) == cdist.INT64_MAX and cdist.cost_to_key(float("inf")) == cdist.INT64_MAX
    assert cdist.key_to_cost(cdist.INT64_MAX) == float("inf")
    assert cdist.gather_best(2.5, 7) == (2.5, 7)            # no process group: identity
