"""small-pathtracer_b200 — B200-native radiance loop for maurock/small-pathtracer.

The product is csrc/ (hand-written sm_100a CUDA kernels behind the C ABI of include/ptb200.h) and
host/ (the C++ mirror of the reference's source-level surface + the `smallpt` executable).
This Python package is plumbing for tests, bench.py and torch.distributed.

The directory name contains a hyphen (it is the project's name), so import it through `_pkg.py`
at the repo root:  `from _pkg import ptb`  (registers the module as `small_pathtracer_b200`).
"""
from . import capi  # noqa: F401
from .capi import *  # noqa: F401,F403
