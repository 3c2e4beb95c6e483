"""Drop-in mirror of the reference's ``model`` package for the DualVGR hot path.

Put this package's parent directory (``dualvgr-videoqa_b200/``) ahead of the reference root on ``sys.path`` and
``import model.models as modelset`` (train.py:20, validate.py:14) resolves here: same constructor, same forward
signature and 7-tuple, same ``state_dict`` keys and shapes — the arithmetic runs in libdualvgr_b200.so.
It is equally importable as ``dualvgr_videoqa_b200.model``."""
import importlib
import importlib.util
import os
import sys


def _load_backend():
    """Returns the product package (``dualvgr_videoqa_b200``) whether this package was imported as its sub-package or
    as the top-level ``model`` package of a reference checkout."""
    if "dualvgr_videoqa_b200" in sys.modules:
        return sys.modules["dualvgr_videoqa_b200"]
    if __name__.startswith("dualvgr_videoqa_b200."):
        return importlib.import_module("dualvgr_videoqa_b200")
    pkg_dir = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("dualvgr_videoqa_b200", os.path.join(pkg_dir, "__init__.py"),
                                                  submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["dualvgr_videoqa_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


_load_backend()
