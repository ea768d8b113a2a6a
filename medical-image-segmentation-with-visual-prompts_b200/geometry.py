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


class RowMap:
    """Pair of device index maps for functional.gather_rows: `fwd` int32 [rows_dst] (dst row <- src row, -1 = zeros)
    and its adjoint `bwd` int32 [rows_src]."""

    def __init__(self, fwd: np.ndarray, bwd: np.ndarray):
        self.fwd_host, self.bwd_host = fwd.astype(np.int32), bwd.astype(np.int32)
        self.rows_dst, self.rows_src = int(fwd.shape[0]), int(bwd.shape[0])
        self._dev = {}

    def on(self, device):
        key = str(device)
        t = self._dev.get(key)
        if t is None:
            t = (torch.from_numpy(self.fwd_host).to(device), torch.from_numpy(self.bwd_host).to(device))
            self._dev[key] = t
        return t


def _invert(slot2vox: np.ndarray, n_vox: int) -> np.ndarray:
    """vox -> slot (-1 where no slot holds the voxel); slot2vox is injective on its non-negative entries."""
    inv = np.full(n_vox, -1, dtype=np.int64)
    ok = slot2vox >= 0
    inv[slot2vox[ok]] = np.nonzero(ok)[0]
    return inv


def _compose(dst2vox: np.ndarray, src2vox: np.ndarray, n_vox: int) -> RowMap:
    """Rows of both arrangements are labelled by the voxel they hold (-1 = padding / cropped)."""
    def one(d2v, s2v):
        inv = _invert(s2v, n_vox)
        out = np.full(d2v.shape[0], -1, dtype=np.int64)
        ok = d2v >= 0
        out[ok] = inv[d2v[ok]]
        return out
    return RowMap(one(dst2vox, src2vox), one(src2vox, dst2vox))


_MERGE_OFFS3 = ((0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (1, 0, 1), (0, 1, 1), (1, 1, 1))   # down.py:31-39
_MERGE_OFFS2 = ((0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 0))                                                 # down.py:41-45


def merge_rows_to_voxels(dims: Sequence[int], merge_last_dim: bool) -> Tuple[np.ndarray, Tuple[int, int, int]]:
    """Voxel held by every row (t', k) of PatchMerging's gathered tensor [T'][K][C] (down.py:21-47): odd axes get
    one zero plane on the LOW side (F.pad with the reversed list, :26-28).  Returns (rows -> voxel | -1, merged dims)."""
    H, W, D = (int(v) for v in dims)
    ph, pw, pd = H % 2, W % 2, D % 2
    h2, w2 = (H + ph) // 2, (W + pw) // 2
    d2 = (D + pd) // 2 if merge_last_dim else D + pd
    offs = _MERGE_OFFS3 if merge_last_dim else _MERGE_OFFS2
    hh, ww, dd, kk = np.meshgrid(np.arange(h2), np.arange(w2), np.arange(d2), np.arange(len(offs)), indexing='ij')
    o = np.asarray(offs)
    h = 2 * hh + o[kk, 0] - ph
    w = 2 * ww + o[kk, 1] - pw
    d = (2 * dd + o[kk, 2] if merge_last_dim else dd) - pd
    ok = (h >= 0) & (w >= 0) & (d >= 0)
    vox = np.where(ok, (h * W + w) * D + d, -1).reshape(-1)
    return vox.astype(np.int64), (h2, w2, d2)


@functools.lru_cache(maxsize=256)
def rowmap_regroup(ga: "Geometry", gb: "Geometry") -> RowMap:
    """block output tokens in window order `ga`  ->  block input tokens in window order `gb` (same feature map)."""
    n_vox = int(np.prod(ga.dims))
    return _compose(gb.index_map_host(0).reshape(-1).astype(np.int64), ga.index_map_host(1).reshape(-1).astype(np.int64), n_vox)


@functools.lru_cache(maxsize=256)
def rowmap_from_voxels(g: "Geometry") -> RowMap:
    """channels-last feature map rows [H*W*D]  ->  block input tokens in window order `g`."""
    n_vox = int(np.prod(g.dims))
    return _compose(g.index_map_host(0).reshape(-1).astype(np.int64), np.arange(n_vox, dtype=np.int64), n_vox)


@functools.lru_cache(maxsize=256)
def rowmap_merge(g: "Geometry", merge_last_dim: bool) -> Tuple[RowMap, Tuple[int, int, int]]:
    """block output tokens in window order `g`  ->  PatchMerging rows [T'*K]."""
    n_vox = int(np.prod(g.dims))
    vox, mdims = merge_rows_to_voxels(g.dims, merge_last_dim)
    return _compose(vox, g.index_map_host(1).reshape(-1).astype(np.int64), n_vox), mdims


@functools.lru_cache(maxsize=256)
def rowmap_merge_from_voxels(dims: Tuple[int, int, int], merge_last_dim: bool) -> Tuple[RowMap, Tuple[int, int, int]]:
    """channels-last feature map rows [H*W*D]  ->  PatchMerging rows [T'*K]."""
    n_vox = int(np.prod(dims))
    vox, mdims = merge_rows_to_voxels(dims, merge_last_dim)
    return _compose(vox, np.arange(n_vox, dtype=np.int64), n_vox), mdims


@functools.lru_cache(maxsize=256)
def get_geometry(dims: Tuple[int, ...], ws: Tuple[int, ...], shift_cfg: Tuple[int, ...]) -> Geometry:
    return Geometry(dims, ws, shift_cfg)
