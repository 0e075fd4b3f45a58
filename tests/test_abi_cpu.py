"""CPU suite: the C-ABI library loads and exports every symbol include/mmtpsm_b200.h declares."""
import os
import re

from conftest import ROOT


def test_library_exports_every_declared_symbol():
    import mmt_psm_b200  # noqa: F401
    from mmt_psm_b200 import lib
    handle = lib.lib()
    header = open(os.path.join(ROOT, "include", "mmtpsm_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(mmt_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(handle, name), "missing export: " + name
    # the ctypes prototypes cover exactly the header
    assert declared == set(lib.exported_symbols()), declared ^ set(lib.exported_symbols())
    assert b"sm_100a" in handle.mmt_version()


def test_no_cpu_fallback():
    import pytest
    import torch
    from mmt_psm_b200 import kernels
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        kernels.nms(torch.zeros(3, 4), torch.zeros(3), 0.5)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        kernels.roi_align_forward(torch.zeros(1, 4, 8, 8), torch.zeros(1, 5), 1.0, 7, 7, 2)
