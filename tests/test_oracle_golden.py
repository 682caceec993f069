"""CPU: the oracle's full render against fixtures produced by the reference itself
(oracle/_ref/smallpt_ref = the reference's src/smallpt.cpp + patches P0-P6; tests/golden/make_golden.py)."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import ptb, orc, ROOT

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "smallpt_ref")


@pytest.mark.parametrize("scene", ["A", "B", "C"])
@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("det", [0, 1])
def test_oracle_render_bit_identical_to_reference(golden_render, scene, mode, det):
    w, h, spp = (int(v) for v in golden_render["meta_whs"])
    sc = ptb.builtin_scene(scene, w, h)
    p = ptb.params(w, h, spp, mode=mode, engine=ptb.PT_ENGINE_FP64_ERAND48, sincos=det)
    cl, mean, sq, st = orc.oracle_render(sc, p)
    assert np.array_equal(cl, golden_render[f"{scene}_{mode}_{det}_clamped"])
    assert np.array_equal(mean, golden_render[f"{scene}_{mode}_{det}_mean"])
    assert np.array_equal(sq, golden_render[f"{scene}_{mode}_{det}_sumsq"])
    assert st.paths == w * h * spp


def test_oracle_is_thread_count_independent():
    # rows own their RNG stream (src/smallpt.cpp:530): OpenMP scheduling must not change the image
    sc = ptb.builtin_scene("A", 40, 30)
    p = ptb.params(40, 30, 4, mode=0, engine=1)
    L = orc.load_oracle()
    n0 = L.oracle_set_threads(0)
    a = orc.oracle_render(sc, p)[0]
    L.oracle_set_threads(1)
    b, st = orc.oracle_render(sc, p)[0], orc.oracle_render(sc, p)[3]
    L.oracle_set_threads(n0)
    assert st.threads == 1
    assert np.array_equal(a, b)


def test_ppm_writer_byte_exact(golden_render, tmp_path):
    w, h, spp = (int(v) for v in golden_render["meta_whs"])
    want = golden_render["A_0_0_ppm"].tobytes()
    assert want.startswith(b"P3\n%d %d\n255\n" % (w, h))
    img = golden_render["A_0_0_clamped"]
    # host C++ writer (the product's) and the oracle's writer both reproduce the reference's bytes (:548-551)
    path = str(tmp_path / "host.ppm")
    ptb.write_ppm(path, img, w, h)
    assert open(path, "rb").read() == want
    L = orc.load_oracle()
    path2 = str(tmp_path / "oracle.ppm")
    a = np.ascontiguousarray(img)
    import ctypes as C
    L.oracle_write_ppm(path2.encode(), a.ctypes.data_as(C.POINTER(C.c_double)), w, h)
    assert open(path2, "rb").read() == want


def test_survey_appendix_f_pixels():
    # oracle P0-P5, 512x512, 16 spp, scene A — dyadic COS/UNI values and the NEE pixel (0,0)
    sc = ptb.builtin_scene("A", 512, 512)
    # only the first row tile is rendered (tile_rows=8, world=64 -> rows 0..7), pixels of row 0 are unaffected
    p = ptb.params(512, 512, 16, mode=0, engine=1, tile_rows=8, rank=0, world=64)
    cl = orc.oracle_render(sc, p)[0]
    assert cl[0, 0].tolist() == [0.14841266228783917, 0.21147312050371372, 0.13991414884031766]
    p = ptb.params(512, 512, 16, mode=2, engine=1, tile_rows=8, rank=0, world=64)
    cl = orc.oracle_render(sc, p)[0]
    assert cl[0, 0].tolist() == [0.10546875, 0.31640625, 0.10546875]


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_vs_live_reference_binary():
    w, h, spp = 48, 40, 6
    with tempfile.TemporaryDirectory() as tmp:
        for scene, mode, det in (("A", 0, 0), ("A", 1, 1), ("B", 2, 0), ("C", 0, 1)):
            prefix = os.path.join(tmp, "r")
            subprocess.check_call([REF_BIN, str(spp), str(mode), scene, str(w), str(h), prefix, str(det), "2"],
                                  stdout=subprocess.DEVNULL)
            want = np.fromfile(prefix + ".clamped.f64").reshape(h, w, 3)
            sc = ptb.builtin_scene(scene, w, h)
            cl = orc.oracle_render(sc, ptb.params(w, h, spp, mode=mode, engine=1, sincos=det))[0]
            assert np.array_equal(cl, want), (scene, mode, det)


def test_fixture_statistics_scene_B():
    # the reference's own saved renders pin image means statistically.
    # image2_32pps_importancesampl.ppm (scene B, cosine, 32 spp): mean RGB (131.3,132.8,109.5)
    # image_32pps_totalrandom.ppm (scene B, uniform weight 1, 32 spp): (102.6,104.0,84.8) => no 2cos factor.
    for mode, want in ((1, (131.3, 132.8, 109.5)), (2, (102.6, 104.0, 84.8))):
        sc = ptb.builtin_scene("B", 128, 128)
        cl = orc.oracle_render(sc, ptb.params(128, 128, 32, mode=mode, engine=1))[0]
        ints = np.vectorize(ptb.to_int)(cl)
        got = ints.reshape(-1, 3).mean(axis=0)
        assert np.all(np.abs(got - np.array(want)) < 1.5), (mode, got)
