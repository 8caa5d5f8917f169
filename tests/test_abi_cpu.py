"""The C-ABI shared library must load on a CPU-only box and export every symbol include/svsr.h declares."""
import ctypes

import pytest

from syncvsr_b200 import _lib


@pytest.fixture(scope="module")
def lib():
    if not _lib.LIB_PATH.exists():
        import __graft_entry__ as g

        g.build()
    return _lib.lib()


def test_every_declared_symbol_is_exported(lib):
    names = _lib.declared_symbols()
    assert len(names) >= 5
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/svsr.h but not exported by libsvsr.so"


def test_abi_version_and_error_string(lib):
    assert lib.svsr_abi_version() >= 1
    assert isinstance(lib.svsr_last_error(), bytes)


def test_argument_validation_needs_no_gpu(lib):
    rc = lib.svsr_gemm_bf16(None, 64, None, 64, None, 64, None, None, 0, 0, 0, 0, 0, ctypes.c_float(1.0), None)
    assert rc == -1
    assert b"empty" in lib.svsr_last_error()
