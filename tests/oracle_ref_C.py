"""Loader of the reference's own compiled CPU `_C` (oracle/_ref/_C*.so, built by oracle/build_ref.py)."""
import glob
import importlib.util
import os

import torch  # noqa: F401  (libtorch must be loaded first)

_here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_cands = glob.glob(os.path.join(_here, "oracle", "_ref", "_C*.so"))
if not _cands:
    raise ImportError("oracle/_ref/_C*.so not built")
_spec = importlib.util.spec_from_file_location("_C", _cands[0])
_mod = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mod)
nms = _mod.nms
roi_align_forward = _mod.roi_align_forward
