import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# hermetic tests: every specialisation is really built by NVRTC (the disk cache has its own test)
os.environ.setdefault("PTB200_CACHE_DIR", "off")

from _pkg import ptb  # noqa: E402
from oracle import pyoracle as orc  # noqa: E402  (the CPU checker's loader: test infrastructure, not part of the product package)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def pt():
    return ptb


@pytest.fixture(scope="session")
def golden_render():
    return np.load(os.path.join(GOLDEN, "ref_render.npz"))


@pytest.fixture(scope="session")
def golden_units():
    return np.load(os.path.join(GOLDEN, "ref_units.npz"))


def room_rays(n, seed, f32_exact=True, margin=0.0):
    """Random rays from inside the built-in room (x 1..99, y 0..81.6, z 0..170)."""
    rng = np.random.default_rng(seed)
    o = np.stack([rng.uniform(1 + margin, 99 - margin, n), rng.uniform(margin, 81.6 - margin, n),
                  rng.uniform(margin, 170 - margin, n)], 1)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    if f32_exact:
        o = o.astype(np.float32).astype(np.float64)
        d = d.astype(np.float32).astype(np.float64)
    return np.ascontiguousarray(np.concatenate([o, d], 1))
