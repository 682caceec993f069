"""GPU: scenes as data.  A scene read from the text format renders the same image, bit for bit, as the built-in table,
and the `smallpt` executable writes the reference's P3 image plus the binary / float / variance side outputs."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ptb, ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,mode", [("A", 0), ("B", 1), ("synthetic", 1)])
def test_scene_file_renders_like_the_builtin_table(name, mode):
    w, h, spp = 96, 64, 16
    imgs = []
    for sc in (ptb.builtin_scene(name, w, h), ptb.parse_scene(ptb.scene_text(name), w, h)):
        with ptb.Context(sc) as c:
            c.render(ptb.params(w, h, spp, mode=mode, seed=5))
            imgs.append(c.readback()[0])
    if name == "synthetic":          # tilted-plane frames differ in the last bit after the text round trip
        assert np.isclose(imgs[0], imgs[1], rtol=1e-6, atol=1e-9).mean() > 0.995
    else:
        assert np.array_equal(imgs[0], imgs[1])


def test_smallpt_executable_outputs(tmp_path):
    exe = os.path.join(ROOT, "small-pathtracer_b200", "smallpt")
    scene = tmp_path / "a.scene"
    out = {k: str(tmp_path / f"img.{k}") for k in ("ppm", "p6", "pfm", "f64", "var")}
    r = subprocess.run([exe, "8", "--scene", "A", "--size", "64x48", "--dump-scene", str(scene), "--out", out["ppm"]], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, "8", "--scene-file", str(scene), "--size", "64x48", "--out", out["ppm"] + "2", "--ppm6", out["p6"], "--pfm", out["pfm"],
                        "--raw64", out["f64"], "--variance", out["var"]], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # NB: --variance switches per-pixel statistics on, which must not change the image (same Philox streams)
    assert open(out["ppm"]).read() == open(out["ppm"] + "2").read()
    toks = open(out["ppm"]).read().split()
    assert toks[:4] == ["P3", "64", "48", "255"] and len(toks) == 4 + 64 * 48 * 3
    p6 = open(out["p6"], "rb").read()
    assert np.array_equal(np.frombuffer(p6[len(b"P6\n64 48\n255\n"):], dtype=np.uint8), np.array(toks[4:], dtype=np.uint8))
    mean = np.frombuffer(open(out["f64"], "rb").read().split(b"\n", 1)[1], dtype="<f8")
    var = np.frombuffer(open(out["var"], "rb").read().split(b"\n", 1)[1], dtype="<f8")
    assert mean.size == var.size == 64 * 48 * 3 and (var >= 0).all() and var.max() > 0 and 0.05 < mean.mean() < 2
    pfm = open(out["pfm"], "rb").read()
    data = np.frombuffer(pfm[len(b"PF\n64 48\n-1.0\n"):], dtype="<f4").reshape(48, 64, 3)[::-1]
    assert np.allclose(data.reshape(-1), mean.astype(np.float32))


def test_smallpt_executable_chunked_checkpoint_resume(tmp_path):
    exe = os.path.join(ROOT, "small-pathtracer_b200", "smallpt")
    a, b, ck = str(tmp_path / "a.ppm"), str(tmp_path / "b.ppm"), str(tmp_path / "ck.f64")
    base = [exe, "--scene", "A", "--size", "64x48", "--seed", "9"]
    assert subprocess.run([exe, "24"] + base[1:] + ["--out", a], capture_output=True).returncode == 0
    # 10 samples in chunks of 4 with a checkpoint, then resume up to 24
    r = subprocess.run([exe, "10"] + base[1:] + ["--chunk", "4", "--checkpoint", ck, "--out", b], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(ck, "rb").readline().split()[:5] == [b"PTB200F64", b"64", b"48", b"3", b"10"]
    r = subprocess.run([exe, "24"] + base[1:] + ["--resume", ck, "--chunk", "5", "--out", b], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(a).read() == open(b).read()
