"""Window geometry of one block call (host logic lives in C: csrc/host.cpp, include/pwa.h).

Mirrors the shape bookkeeping at the top of the reference's
SwinTransformerBlock.forward_attn_mlp (swin_transformer/swin_block.py:146-164, 265-270) and
get_attn_mask (:312-364) in compressed form (one uint8 region id per window token).
"""
from __future__ import annotations

import ctypes as C
import functools
from typing import Sequence, Tuple

import numpy as np
import torch

from . import _lib


class Geometry:
    def __init__(self, dims: Sequence[int], ws: Sequence[int], shift_cfg: Sequence[int]):
        self.c = _lib.PwaGeom()
        a3 = C.c_int32 * 3
        rc = _lib.lib.pwa_geometry(a3(*map(int, dims)), a3(*map(int, ws)), a3(*map(int, shift_cfg)), C.byref(self.c))
        _lib.check(rc, "pwa_geometry")
        g = self.c
        self.dims: Tuple[int, ...] = tuple(g.dims)
        self.ws: Tuple[int, ...] = tuple(g.ws)
        self.shift: Tuple[int, ...] = tuple(g.shift)
        self.pads: Tuple[int, ...] = tuple(g.pads)
        self.sp: Tuple[int, ...] = tuple(g.sp)
        self.nwin: Tuple[int, ...] = tuple(g.nwin)
        self.P, self.N = int(g.P), int(g.N)
        self.masked, self.padded = bool(g.masked), bool(g.padded)
        self._ids_host = None
        self._ids_dev = {}

    def ref(self):
        return C.byref(self.c)

    def region_ids_host(self) -> np.ndarray:
        """uint8 [P, N]; mask[p,i,j] = (ids[p,i] == ids[p,j])."""
        if self._ids_host is None:
            ids = np.empty((self.P, self.N), dtype=np.uint8)
            _lib.check(_lib.lib.pwa_region_ids(self.ref(), ids.ctypes.data), "pwa_region_ids")
            self._ids_host = ids
        return self._ids_host

    def region_ids(self, device) -> torch.Tensor:
        """Device copy, uploaded once per (geometry, device) and cached."""
        key = str(device)
        t = self._ids_dev.get(key)
        if t is None:
            t = torch.from_numpy(self.region_ids_host()).to(device)
            self._ids_dev[key] = t
        return t

    def index_map_host(self, which: int) -> np.ndarray:
        m = np.empty((self.P, self.N), dtype=np.int32)
        _lib.check(_lib.lib.pwa_index_map(self.ref(), int(which), m.ctypes.data), "pwa_index_map")
        return m


@functools.lru_cache(maxsize=256)
def get_geometry(dims: Tuple[int, ...], ws: Tuple[int, ...], shift_cfg: Tuple[int, ...]) -> Geometry:
    return Geometry(dims, ws, shift_cfg)
