"""ctypes binding of libsvsr.so (the C ABI in include/svsr.h).

The product path has no fallback: if the shared object is missing or a call fails, this raises."""
from __future__ import annotations

import ctypes
import re
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libsvsr.so"
HEADER_PATH = _PKG.parent / "include" / "svsr.h"

_lib: ctypes.CDLL | None = None


class SvsrError(RuntimeError):
    pass


def declared_symbols() -> list[str]:
    """Every function name declared in include/svsr.h (used by the CPU-side ABI test)."""
    text = HEADER_PATH.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(svsr_[a-z0-9_]+)\s*\(", text)))


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise SvsrError(
                f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU or eager fallback for the hot path)"
            )
        _lib = ctypes.CDLL(str(LIB_PATH))
        _lib.svsr_last_error.restype = ctypes.c_char_p
        _lib.svsr_abi_version.restype = ctypes.c_int
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().svsr_last_error().decode(errors="replace")
        raise SvsrError(f"{what} failed (status {rc}): {msg}")


def ptr(t) -> ctypes.c_void_p:
    """Device pointer of a torch tensor (or None)."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr() -> ctypes.c_void_p:
    import torch

    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
