"""CPU: the data formats either side of the hot path (host/scene_io.hpp): the text scene format
round-trips the built-in tables, the binary/float image writers agree with the reference's P3 writer (:548-551)."""
import ctypes as C
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import ptb, ROOT


def _same_struct(a, b):
    return bytes(a) == bytes(b)


@pytest.mark.parametrize("name", ["A", "B", "C"])
def test_text_format_round_trips_the_builtin_tables_bit_for_bit(name):
    ref = ptb.builtin_scene(name, 320, 200)
    text = ptb.scene_text(name)
    got = ptb.parse_scene(text, 320, 200)
    assert (got.n_spheres, got.n_planes) == (ref.n_spheres, ref.n_planes)
    assert list(got.order) == list(ref.order)
    assert all(_same_struct(got.spheres[i], ref.spheres[i]) for i in range(ref.n_spheres))
    assert all(_same_struct(got.planes[i], ref.planes[i]) for i in range(ref.n_planes))
    assert _same_struct(got.light, ref.light)
    assert _same_struct(got.camera, ref.camera)            # `camera` statement = the literals of :65,:521


def test_synthetic_scene_round_trip_keeps_tilted_frames():
    ref = ptb.builtin_scene("synthetic", 64, 64)
    got = ptb.parse_scene(ptb.scene_text("synthetic"), 64, 64)
    assert got.n_planes == ref.n_planes and got.n_spheres == ref.n_spheres
    for i in range(ref.n_planes):
        a, b = got.planes[i], ref.planes[i]
        assert a.kind == b.kind and a.refl == b.refl
        for f in ("p0", "n", "s", "t"):                   # the Plane ctor re-normalises: last-bit differences only
            assert np.allclose(getattr(a, f).tup(), getattr(b, f).tup(), rtol=0, atol=1e-14)


def test_parser_reports_errors_with_line_numbers():
    with pytest.raises(ptb.PtError, match="line 2.*unknown statement"):
        ptb.parse_scene("sphere 1 0 0 0 0 0 0 1 1 1 DIFF\ncube 1 2 3\n")
    with pytest.raises(ptb.PtError, match="line 1.*needs more numbers"):
        ptb.parse_scene("rect_xz 1 2 3\n")
    with pytest.raises(ptb.PtError, match="material"):
        ptb.parse_scene("sphere 1 0 0 0 0 0 0 1 1 1 GLASS\n")
    with pytest.raises(ptb.PtError, match="no objects"):
        ptb.parse_scene("# nothing here\n")
    sc = ptb.parse_scene("# one sphere\nsphere 2.5 1 2 3  0 0 0  .5 .5 .5 REFR   # trailing comment\n")
    assert sc.n_spheres == 1 and sc.spheres[0].refl == ptb.PT_REFR and sc.spheres[0].rad == 2.5 and sc.light.id == -1


def test_binary_ppm_and_pfm_match_the_p3_writer(tmp_path):
    rng = np.random.default_rng(3)
    w, h = 13, 7
    img = rng.uniform(-0.2, 1.4, size=(h, w, 3))
    p3, p6, pfm, raw = (str(tmp_path / n) for n in ("a.ppm", "b.ppm", "c.pfm", "d.f64"))
    ptb.write_image(p3, img, "ppm"); ptb.write_image(p6, img, "ppm6"); ptb.write_image(pfm, img, "pfm"); ptb.write_image(raw, img, "raw64")
    toks = open(p3).read().split()
    assert toks[:4] == ["P3", str(w), str(h), "255"]
    vals3 = np.array(toks[4:], dtype=np.int64).reshape(h, w, 3)
    b = open(p6, "rb").read()
    header = f"P6\n{w} {h}\n255\n".encode()
    assert b.startswith(header) and len(b) == len(header) + w * h * 3
    vals6 = np.frombuffer(b[len(header):], dtype=np.uint8).reshape(h, w, 3)
    assert np.array_equal(vals3, vals6)                                       # same clamp + gamma (:314-321), bytes instead of text
    f = open(pfm, "rb").read()
    hdr = f"PF\n{w} {h}\n-1.0\n".encode()
    assert f.startswith(hdr)
    data = np.frombuffer(f[len(hdr):], dtype="<f4").reshape(h, w, 3)[::-1]    # PFM rows run bottom to top
    assert np.array_equal(data, img.astype(np.float32))                       # linear, unclamped
    r = open(raw, "rb").read()
    line, body = r.split(b"\n", 1)
    assert line.split()[:4] == [b"PTB200F64", str(w).encode(), str(h).encode(), b"3"]
    assert np.array_equal(np.frombuffer(body, dtype="<f8").reshape(h, w, 3), img)


def test_smallpt_executable_rejects_a_bad_scene_file(tmp_path):
    exe = os.path.join(ROOT, "small-pathtracer_b200", "smallpt")
    bad = tmp_path / "bad.scene"
    bad.write_text("torus 1 2 3\n")
    r = subprocess.run([exe, "4", "--scene-file", str(bad), "--size", "8x8", "--out", str(tmp_path / "x.ppm")], capture_output=True, text=True)
    assert r.returncode == 1 and "unknown statement 'torus'" in r.stderr
