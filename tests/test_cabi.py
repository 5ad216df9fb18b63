"""The C-ABI library loads and exports every symbol include/deepdish_b200.h declares (no GPU needed:
nothing here launches a kernel), and the host-only layout arithmetic is sane."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    from deepdish_b200 import _lib, build
    build.build()            # no-op when up to date; nvcc cross-compiles without a GPU
    return _lib, _lib.lib()


def test_exports_every_declared_symbol():
    _l, lib = _lib()
    hdr = open(os.path.join(ROOT, "include", "deepdish_b200.h")).read()
    names = re.findall(r"^(?:int|const char\*)\s+(dd_\w+)\s*\(", hdr, flags=re.M)
    assert len(names) >= 24
    for n in names:
        assert hasattr(lib, n), n
    assert b"sm_100a" in lib.dd_version()


def test_layout_query_and_invalid_config():
    _l, lib = _lib()
    cfg = _l.make_config(1024, 128, 64, 100, ["person", "bicycle", "car"], max_age=60)
    lay = _l.TrackerLayout()
    assert lib.dd_tracker_layout_query(ctypes.byref(cfg), ctypes.byref(lay)) == 0
    specs = _l.field_specs(cfg)
    offs = sorted((getattr(lay, n), n) for n in _l.LAYOUT_FIELDS)
    for (o, n), (o2, _) in zip(offs, offs[1:] + [(lay.total_bytes, None)]):
        import numpy as np
        dt, shape = specs[n]
        assert o % 256 == 0 and o + int(np.prod(shape)) * np.dtype(dt).itemsize <= o2, n
    # galleries live in the page pool, not in the blob: the blob holds 7-entry page tables for budget 100
    assert specs["ptab"][1] == (1024, 128, 7) and lay.total_bytes < 1024 * 128 * 100 * 512 // 16
    unb = _l.make_config(4, 16, 16, None, ["person"], page_cap=40)          # nn_budget=None (deepdish.py:515)
    assert unb.budget == 0 and _l.field_specs(unb)["ptab"][1] == (4, 16, 40)
    assert lib.dd_tracker_layout_query(ctypes.byref(unb), ctypes.byref(lay)) == 0
    cfg.feat_dim = 64
    assert lib.dd_tracker_layout_query(ctypes.byref(cfg), ctypes.byref(lay)) == _l.DD_ERR_INVALID
    with pytest.raises(ValueError):
        _l.make_config(4, 16, 16, 0, ["person"])
    bad = _l.make_config(4, 16, 16, 10, ["person"], max_age=40000)           # tsu is kept in 16 bits on the device
    assert lib.dd_tracker_layout_query(ctypes.byref(bad), ctypes.byref(lay)) == _l.DD_ERR_INVALID
    assert cfg.label_rank[0] == 2 and cfg.label_rank[1] == 0 and cfg.label_rank[2] == 1


def test_engine_rejects_bad_arguments_without_touching_cuda():
    _l, lib = _lib()
    h = ctypes.c_void_p()
    assert lib.dd_engine_create(0, None, None, None, None, None, None, 0, None, None, None, 0, 0, ctypes.byref(h)) == _l.DD_ERR_INVALID
    assert lib.dd_engine_step(None, None, None, None, None, None, 0, None) == _l.DD_ERR_INVALID
    assert lib.dd_engine_destroy(None) == _l.DD_ERR_INVALID


def test_product_never_imports_oracle():
    """The product package must not reference oracle/ (no CPU fallback path)."""
    pkg = os.path.join(ROOT, "deepdish_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
