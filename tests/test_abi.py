"""CPU-side checks of the C-ABI shared library: it loads and exports every symbol include/wiski_b200.h declares."""
import ctypes
import os
import re

from online_gp_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "wiski_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wiski_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "build the library first: online_gp_b200/csrc/build.sh"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/wiski_b200.h but not exported"


def test_binding_covers_header():
    assert sorted(_lib.EXPORTED) == _declared_symbols()


def test_load_and_version():
    lib = _lib.load()
    assert lib.wiski_abi_version() == 1
    assert lib.wiski_gram_work_elems(1 << 20, 512, 512) > 0
    assert lib.wiski_qmv_work_elems(1 << 20, 512, 1) > 0


def test_ops_refuse_cpu_tensors():
    import pytest
    import torch
    from online_gp_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.gram(torch.zeros(8, 4), torch.zeros(8, 4))
