"""CPU: the C-ABI library loads, exports every symbol include/ptb200.h declares, and fails LOUDLY without a
GPU (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import pytest

from conftest import ptb, ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ptb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pt_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    lib = C.CDLL(ptb.LIB_PATH)
    declared = _declared_symbols()
    assert {"pt_scene_upload", "pt_render", "pt_readback"} <= set(declared)     # the north star's three names
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ptb200.h but not exported"
    assert sorted(ptb.EXPORTS) == declared


def test_struct_layouts_match_header():
    # sizes the C compiler gives the header's PODs (computed by a tiny C program at build time would be
    # circular; these are the natural-alignment sizes of the declarations)
    assert C.sizeof(ptb.Vec3) == 24
    assert C.sizeof(ptb.Sphere) == 8 + 72 + 8
    assert C.sizeof(ptb.Plane) == 8 + 40 + 96 + 16 + 48
    assert C.sizeof(ptb.Camera) == 96
    assert C.sizeof(ptb.Light) == 8 + 48
    assert C.sizeof(ptb.RenderParams) == 24 + 8 + 16 + 16 + 8 + 8
    assert C.sizeof(ptb.Stats) == 72 + 8 + 16 + 8


def test_version_string():
    assert b"sm_100a" in ptb.lib().pt_version()


def test_no_device_is_an_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device path cannot be exercised")
    sc = ptb.builtin_scene("A")
    with pytest.raises(ptb.PtError, match="no CUDA device"):
        ptb.Context(sc)


def test_product_code_never_touches_the_oracle():
    # the product (csrc/, host/, include/) must not reference oracle/ in any way
    bad = []
    for sub in ("small-pathtracer_b200/csrc", "small-pathtracer_b200/host", "include"):
        for fn in os.listdir(os.path.join(ROOT, sub)):
            text = open(os.path.join(ROOT, sub, fn), errors="ignore").read()
            if re.search(r"oracle[/_]|liboracle", text) and fn not in ("ptb200_detmath.h",):
                bad.append(f"{sub}/{fn}")
    assert not bad, bad
