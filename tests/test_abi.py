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
