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


def test_struct_layouts_match_header(tmp_path):
    """sizeof / offsetof of every POD, as the C compiler sees include/ptb200.h, against the ctypes mirrors."""
    import subprocess
    pairs = {"pt_vec3": ptb.Vec3, "pt_sphere": ptb.Sphere, "pt_plane": ptb.Plane, "pt_camera": ptb.Camera, "pt_light": ptb.Light,
             "pt_scene": ptb.SceneDesc, "pt_render_params": ptb.RenderParams, "pt_stats": ptb.Stats}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "ptb200.h"', 'int main(void) {']
    for cname, cls in pairs.items():
        lines.append(f'  printf("{cname} sizeof %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = {}
    for ln in subprocess.check_output([str(exe)], text=True).splitlines():
        c, f, v = ln.split()
        got[(c, f)] = int(v)
    for cname, cls in pairs.items():
        assert got[(cname, "sizeof")] == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert got[(cname, fname)] == getattr(cls, fname).offset, (cname, fname)
    # every field the header declares has a mirror (same count: a field added on one side only fails here)
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "ptb200.h")).read(), flags=re.S)
    for cname, cls in pairs.items():
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), text, flags=re.S).group(1)
        n_decl = sum(len(stmt.split(",")) for stmt in body.split(";") if stmt.strip())
        assert n_decl == len(cls._fields_), (cname, n_decl, len(cls._fields_))


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
