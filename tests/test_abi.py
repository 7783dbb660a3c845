"""CPU: the C-ABI library loads and exports every symbol include/ctrb200.h declares;
compute entry points fail loudly without a GPU (no CPU fallback)."""
import re

import pytest

from conftest import ROOT


def test_exports_match_header():
    from ctrdapp_b200 import _lib
    header = (ROOT / "include" / "ctrb200.h").read_text()
    declared = set(re.findall(r"\b(ctrb_[a-z0-9_]+)\s*\(", header))
    declared = {d for d in declared if not d.endswith("_t")}
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    h = _lib.lib()
    for name in declared:
        assert hasattr(h, name)
    assert b"sm_100a" in h.ctrb_version()


def test_no_cpu_fallback():
    import torch
    from ctrdapp_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert _lib.lib().ctrb_device_count() == 0
    with pytest.raises(_lib.CtrbError, match="no CPU fallback"):
        _lib.Context(0)


def test_product_does_not_import_oracle():
    """The product path may not route through oracle/ (only tests, smoke and bench's cpu legs may)."""
    pkg = ROOT / "ctr-design-and-path-plan_b200"
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")):
        txt = p.read_text()
        assert "oracle" not in txt.lower(), p


def test_library_reads_nothing_from_the_environment():
    """include/ctrb200.h promises immutable handles and per-ctx options: no kernel-selection switch may be read from the
    process environment (round 1 had five getenv calls on the launch paths)."""
    csrc = ROOT / "ctr-design-and-path-plan_b200" / "csrc"
    for p in list(csrc.glob("*.cu")) + list(csrc.glob("*.cuh")):
        assert "getenv" not in p.read_text(), p


@pytest.mark.gpu
def test_ctx_options_round_trip():
    """ctrb_ctx_set_option / ctrb_ctx_get_option: defaults, round trip, range checks (ValueError), restore on exit."""
    from ctrdapp_b200 import _lib
    ctx = _lib.Context(0)
    defaults = {"static_solver": 0, "static_one_step": 0, "fast_generic_n": 0, "fast_block_invert": 1, "eval_group": 1,
                "gcurve_blocks": 3, "max_nodes_hint": 0, "static_curves_scan": 0}
    assert {k: ctx.get_option(k) for k in defaults} == defaults
    with ctx.options(static_one_step=1, gcurve_blocks=5):
        assert ctx.get_option("static_one_step") == 1 and ctx.get_option("gcurve_blocks") == 5
    assert {k: ctx.get_option(k) for k in defaults} == defaults
    with pytest.raises(ValueError):
        ctx.set_option("eval_group", 2)
    with pytest.raises(ValueError):
        ctx.set_option("gcurve_blocks", 0)
    with pytest.raises(ValueError):
        _lib.check(_lib.lib().ctrb_ctx_set_option(ctx.handle, 99, 0))
