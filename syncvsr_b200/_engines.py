"""Per-geometry native step executors kept alive side by side.

An engine (csrc/engine.cu, csrc/engine_lrs.cu) is built for one clip geometry (B, T, H, W): its workspace layout
depends on it. The reference's loops change geometry all the time -- LRW alternates train / val batch shapes, the LRS
datamodule pads every batch to its own longest clip (LRS/video/datamodule/data_module.py:12-43) -- so the modules keep
the engines they have already built (handle + workspace + packed bf16 weights) in an LRU cache instead of destroying and
rebuilding one engine per shape change. Switching back to a cached geometry costs nothing but a weight repack when the
parameters changed in between; CUDA graphs captured against a cached engine stay valid while it is alive."""
from __future__ import annotations

import ctypes as C
import os
from collections import OrderedDict
from typing import Callable, Optional

import torch

from ._lib import check, lib


class Engine:
    __slots__ = ("h", "ws", "ws_bytes", "id", "key", "packed_version", "packed_native", "dev_seed")

    def __init__(self, h, ws, ws_bytes, eid, key):
        self.h, self.ws, self.ws_bytes, self.id, self.key = h, ws, ws_bytes, eid, key
        self.packed_version = -1  # autograd version of the parameter arena at the last repack of THIS engine
        self.packed_native = -1   # native-update counter (mark_weights_updated) at the last repack
        self.dev_seed = False     # LRS: the engine reads its dropout step seed from device memory (svsr_lrs_step_control)


class EngineCache:
    def __init__(self, prefix: str, device: torch.device, max_bytes: Optional[int] = None):
        self.prefix, self.device = prefix, device
        gb = float(os.environ.get("SVSR_ENGINE_CACHE_GB", "48"))
        self.max_bytes = int(gb * (1 << 30)) if max_bytes is None else max_bytes
        self.entries: "OrderedDict[tuple, Engine]" = OrderedDict()
        self._next_id = 1
        L = lib()
        getattr(L, f"{prefix}_workspace_bytes").restype = C.c_int64

    def alive(self, eid: int) -> bool:
        return any(e.id == eid for e in self.entries.values())

    def get(self, key: tuple) -> Optional[Engine]:
        ent = self.entries.get(key)
        if ent is not None:
            self.entries.move_to_end(key)
        return ent

    def create(self, key: tuple, cfg, bind: Callable[[C.c_void_p, int, int], None],
               before_bind: Optional[Callable[[C.c_void_p], None]] = None) -> Engine:
        """cfg: the ctypes config struct; bind(handle, ws_ptr, ws_bytes) attaches the arenas + workspace."""
        L = lib()
        h = C.c_void_p()
        check(getattr(L, f"{self.prefix}_create")(C.byref(cfg), C.byref(h)), f"{self.prefix}_create")
        if before_bind is not None:
            before_bind(h)  # the first engine of a module: the parameter arenas are created from its layout
        ws_bytes = int(getattr(L, f"{self.prefix}_workspace_bytes")(h))
        self._evict(ws_bytes)
        ws = torch.empty(ws_bytes + 1024, dtype=torch.uint8, device=self.device)
        bind(h, (ws.data_ptr() + 1023) & ~1023, ws_bytes)
        ent = Engine(h, ws, ws_bytes, self._next_id, key)
        self._next_id += 1
        self.entries[key] = ent
        return ent

    def _evict(self, incoming: int) -> None:
        total = sum(e.ws_bytes for e in self.entries.values()) + incoming
        while total > self.max_bytes and self.entries:
            _, old = self.entries.popitem(last=False)  # least recently used
            getattr(lib(), f"{self.prefix}_destroy")(old.h)
            total -= old.ws_bytes
            old.ws = None

    def destroy(self) -> None:
        for e in self.entries.values():
            try:
                getattr(lib(), f"{self.prefix}_destroy")(e.h)
            except Exception:
                pass
        self.entries.clear()
