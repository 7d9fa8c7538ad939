"""CPU: the C-ABI library builds, loads and exports every symbol include/cnl_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    import __graft_entry__ as g
    g.build()
    from centernet_lightning_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH)
    return _lib.LIB_PATH


def _declared():
    src = open(os.path.join(ROOT, "include", "cnl_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cnl_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib_path):
    lib = ctypes.CDLL(lib_path)
    names = _declared()
    assert "cnl_decode_detections" in names and "cnl_engine_forward" in names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/cnl_b200.h but not exported"


def test_binding_lists_every_export(lib_path):
    from centernet_lightning_b200 import _lib
    assert sorted(_lib.EXPORTS) == _declared()
    lib = _lib.load()
    assert lib.cnl_compiled_sm() == 100
    assert lib.cnl_version() >= 1000
    assert lib.cnl_decode_workspace_bytes(32, 128, 128) >= 32 * 128 * 128 * 6


def test_argument_errors_without_gpu(lib_path):
    """Argument validation happens before any CUDA call, so it is checkable on CPU."""
    from centernet_lightning_b200 import _lib
    lib = _lib.load()
    buf = (ctypes.c_char * 4096)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    args = dict(n=1, c=2, h=4, w=4)
    st = lib.cnl_decode_detections(p, p, None, 1, 2, 4, 4, 0, 0, 4, 5, 0, 0, 1.0, 4, p, p, p, p, None, p, 4096, None)
    assert st == 1 and b"odd" in lib.cnl_last_error()
    st = lib.cnl_decode_detections(p, p, None, 1, 2, 4, 4, 0, 0, 3, 17, 0, 0, 1.0, 4, p, p, p, p, None, p, 4096, None)
    assert st == 1 and b"num_detections" in lib.cnl_last_error()
    st = lib.cnl_decode_detections(None, p, None, 1, 2, 4, 4, 0, 0, 3, 5, 0, 0, 1.0, 4, p, p, p, p, None, p, 4096, None)
    assert st == 1
    with pytest.raises(ValueError):
        _lib.check(st, "decode")


def test_cpu_tensors_are_refused():
    import torch
    from centernet_lightning_b200 import decode
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        decode.decode_detections(torch.rand(1, 2, 8, 8), torch.rand(1, 4, 8, 8), num_detections=4)
