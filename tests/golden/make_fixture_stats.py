#!/usr/bin/env python3
"""Statistics of the reference's own saved renders (/root/reference/*.ppm: P3, 512x512, maxval 255).

No fixture is bit-reproducible (the jitter came from a time-seeded libc rand() on another libc), so they serve as
statistical known answers: the image mean in 8-bit gamma space and 16x16-block means.  This script reads the PPMs where
they lie and writes tests/golden/fixture_stats.json; tests/test_fixture_stats.py re-derives the numbers whenever
/root/reference is present (so the committed file is never hand-typed) and the GPU tests compare renders with them."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PTB200_REFERENCE", "/root/reference")

# fixture -> what it shows: (scene, mode, spp, pinned?)
FIXTURES = {
    "image1_16ssp_importsampl.ppm": ("B", "cos", 16, True),
    "image2_32pps_importancesampl.ppm": ("B", "cos", 32, True),
    "image_32pps_totalrandom.ppm": ("B", "uni", 32, True),
    "image_512pps_explicitlight_test.ppm": ("B", "cos", 16, True),      # mis-named: statistically the 16-spp cosine image
    # rect walls + rect light + the two spheres of :297-298, NEE of HEAD; spp is not recorded anywhere: the source's default
    # samps = 16 (:508) reproduces the fixture's 8-bit mean to 0.03 (4/8/32/64 spp are 3.3/1.5/0.8/1.3 away)
    "image_light_test.ppm": ("C", "nee", 16, True),
    "image2_16ssp_explicitsampling.ppm": ("B", "sphere-era NEE (source not in the repo)", 16, False),
    "image2_32pps_explicitsampling.ppm": ("B", "sphere-era NEE (source not in the repo)", 32, False),
    "image_512pps_explicitlight.ppm": ("B", "sphere-era NEE (source not in the repo)", 512, False),
    "image_512pps_random_test.ppm": ("B", "sphere-era NEE (source not in the repo)", 512, False),
    "image_32pps_halflighthalfimportance.ppm": ("B", "sphere-era NEE (source not in the repo)", 32, False),
}


def read_p3(path):
    tok = open(path).read().split()
    assert tok[0] == "P3", path
    w, h, maxval = int(tok[1]), int(tok[2]), int(tok[3])
    a = np.array(tok[4:4 + w * h * 3], dtype=np.int32).reshape(h, w, 3)
    return a, maxval


def stats():
    out = {}
    for name, (scene, mode, spp, pinned) in FIXTURES.items():
        a, maxval = read_p3(os.path.join(REF, name))
        h, w, _ = a.shape
        lin = (a / 255.0) ** 2.2                                   # back to (clamped) linear means
        blocks = lin.reshape(h // 32, 32, w // 32, 32, 3).mean(axis=(1, 3))
        out[name] = {"scene": scene, "mode": mode, "spp": spp, "pinned": pinned, "width": w, "height": h, "maxval": maxval,
                     "mean_rgb_8bit": [round(float(x), 4) for x in a.reshape(-1, 3).mean(axis=0)],
                     "mean_rgb_linear": [round(float(x), 6) for x in lin.reshape(-1, 3).mean(axis=0)],
                     "block_means_linear_16x16": np.round(blocks, 5).tolist()}
    return out


if __name__ == "__main__":
    if not os.path.isdir(REF):
        raise SystemExit(f"{REF} not present")
    with open(os.path.join(HERE, "fixture_stats.json"), "w") as f:
        json.dump(stats(), f)
    print("wrote fixture_stats.json")
    sys.exit(0)
