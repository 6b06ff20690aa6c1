"""CPU-side checks of the boundary: the C-ABI library builds, loads and exports every symbol the
header declares; the ctypes struct mirrors match the C layout; the host-side grid tables of the
product equal the oracle's (bit-exact hash indices depend on it).  No compute calls (no GPU)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from dns_slam_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    L = _lib.lib()                      # also asserts the struct sizes
    hdr = open(os.path.join(ROOT, "include", "dns_slam_b200.h")).read()
    declared = set(re.findall(r"\b(dns_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SYMBOLS)
    assert L.dns_version() >= 100


def test_no_cpu_fallback():
    import torch
    from dns_slam_b200 import tcnn
    enc = tcnn.Encoding(3, {"otype": "OneBlob", "n_bins": 16}, device="cpu")
    with pytest.raises(RuntimeError):
        enc(torch.rand(4, 3))


def test_grid_tables_equal_oracle():
    from dns_slam_b200 import grid
    from oracle.tcnn_standin import grid_level_tables
    for res, hs in ((592, 16), (231, 20), (124, 13)):
        pls = grid.per_level_scale(res)
        a = grid.level_tables(16, 16, pls, hs)
        b = grid_level_tables(16, 16, pls, hs)
        assert a["res"] == [int(v) for v in b["res"]] and a["size"] == [int(v) for v in b["size"]]
        assert a["offset"] == [int(v) for v in b["offset"]] and a["hashed"] == [int(v) for v in b["hashed"]]
        assert np.array_equal(np.asarray(a["scale"], np.float32), b["scale"])


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "dns_slam_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn
