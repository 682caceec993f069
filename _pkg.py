"""Loader for the hyphen-named package directory `small-pathtracer_b200/` (module alias
`small_pathtracer_b200`).  `from _pkg import ptb`."""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
_NAME = "small_pathtracer_b200"

if _NAME in sys.modules:
    ptb = sys.modules[_NAME]
else:
    _spec = importlib.util.spec_from_file_location(
        _NAME, os.path.join(_ROOT, "small-pathtracer_b200", "__init__.py"),
        submodule_search_locations=[os.path.join(_ROOT, "small-pathtracer_b200")])
    ptb = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = ptb
    _spec.loader.exec_module(ptb)
