"""CPU checks of the C-ABI boundary: the library builds, loads, and exports every symbol declared in
include/lpic_b200.h with a ctypes signature; host memory helpers work without a GPU; creating a context
without a CUDA device fails loudly (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

from lambdapic_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.lib()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "lpic_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lpic_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(L):
    names = declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/lpic_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert set(_lib.SIGNATURES) <= set(names)


def test_version_and_host_memory(L):
    assert b"sm_100a" in L.lpic_version()
    from lambdapic_b200.engine import HostBuffer
    b = HostBuffer(1024)
    a = b.array(np.float64, 128)
    a[:] = 3.0
    assert a.sum() == 384.0
    b.free()


def test_no_cpu_fallback(L):
    if L.lpic_device_count() > 0:
        pytest.skip("a GPU is present")
    ctx = L.lpic_create(3, 1, 4, 4, 4, 3, 1.0, 1.0, 1.0, 1, 0)
    assert not ctx
    assert b"no CPU fallback" in L.lpic_last_error()
    from lambdapic_b200.engine import DeviceEngine
    with pytest.raises(_lib.LpicError):
        DeviceEngine(3, 1, 4, 4, 4, 3, 1.0, 1.0, 1.0, 1)
