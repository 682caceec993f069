"""CPU: the reference's own saved renders as statistical known answers.

tests/golden/fixture_stats.json holds their image means; whenever /root/reference is present the numbers are re-derived
from the PPM files themselves (the committed file is never hand-typed), and the oracle must reproduce the pinned ones
through the reference's own output formula (clamp the per-pixel mean, toInt: src/smallpt.cpp:314-321, :538)."""
import json
import os
import sys

import numpy as np
import pytest

from conftest import ptb, orc, GOLDEN

MODE = {"nee": 0, "cos": 1, "uni": 2}


def load_stats():
    with open(os.path.join(GOLDEN, "fixture_stats.json")) as f:
        return json.load(f)


def to_8bit(mean):
    return np.floor(np.clip(mean, 0.0, 1.0) ** (1 / 2.2) * 255 + .5)


def test_committed_statistics_are_the_reference_files():
    ref = os.environ.get("PTB200_REFERENCE", "/root/reference")
    if not os.path.isdir(ref):
        pytest.skip("the reference tree is not on this box")
    sys.path.insert(0, GOLDEN)
    import make_fixture_stats as m
    assert m.stats() == load_stats()


@pytest.mark.parametrize("name", ["image1_16ssp_importsampl.ppm", "image_32pps_totalrandom.ppm", "image_light_test.ppm"])
def test_oracle_reproduces_pinned_fixture_means(name):
    fx = load_stats()[name]
    assert fx["pinned"]
    w, h = fx["width"], fx["height"]
    sc = ptb.builtin_scene(fx["scene"], w, h)
    mean = orc.oracle_render(sc, ptb.params(w, h, fx["spp"], mode=MODE[fx["mode"]], engine=1))[1]
    got = to_8bit(mean).reshape(-1, 3).mean(axis=0)
    assert np.all(np.abs(got - np.array(fx["mean_rgb_8bit"])) < 0.75), (name, got, fx["mean_rgb_8bit"])
    # block means (32x32-pixel blocks, linear): within the noise of two independent renders at this spp
    lin = (to_8bit(mean) / 255.0) ** 2.2
    blocks = lin.reshape(h // 32, 32, w // 32, 32, 3).mean(axis=(1, 3))
    want = np.array(fx["block_means_linear_16x16"])
    rel = np.abs(blocks - want).sum() / want.sum()
    assert rel < 0.08, (name, rel)
