"""Host logic of the token pipeline (no GPU): the row maps that csrc/gather.cu consumes are composed on the host from
pwa_index_map(); here they are applied with numpy and compared with the oracle's reverse -> (pad/roll) partition chain
and with the reference's PatchMerging gather (F.pad + strided slices + cat, down.py:21-47)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import pwa_b200
from pwa_b200 import geometry as G
from oracle import restatement as R

GEOMS = [((8, 8, 4), (4, 4, 2)), ((6, 6, 6), (4, 4, 2)), ((5, 9, 3), (4, 4, 2)), ((12, 12, 24), (8, 8, 4)),
         ((16, 16, 8), (8, 8, 4)), ((6, 8, 4), (4, 4, 2)), ((8, 8, 2), (4, 4, 2))]


def _apply(rows, idx):
    """numpy model of pwa_gather_rows: rows [B, R, C], idx [R'] -> [B, R', C]."""
    out = np.zeros((rows.shape[0], idx.shape[0], rows.shape[2]), dtype=rows.dtype)
    ok = idx >= 0
    out[:, ok] = rows[:, idx[ok]]
    return out


def _geoms(dims, ws):
    shift = tuple(w // 2 for w in ws)
    return pwa_b200.get_geometry(dims, ws, (0, 0, 0)), pwa_b200.get_geometry(dims, ws, shift)


@pytest.mark.parametrize("dims,ws", GEOMS)
def test_regroup_map_equals_reverse_then_partition(dims, ws):
    g0, g1 = _geoms(dims, ws)
    rm = G.rowmap_regroup(g0, g1)
    gen = torch.Generator().manual_seed(3)
    t0 = torch.randn(2, g0.P, g0.N, 5, generator=gen)
    pads = R.pad_amounts(dims, ws)
    x = R.reverse_tokens(t0, dims, ws, R.effective_shift(dims, ws, (0, 0, 0)), pads)
    exp = R.partition_tokens(x, ws, R.effective_shift(dims, ws, tuple(w // 2 for w in ws)), pads)
    got = _apply(t0.reshape(2, -1, 5).numpy(), rm.fwd_host)
    assert np.array_equal(got, exp.reshape(2, -1, 5).numpy())
    # adjoint: <gather(a), g> == <a, gather_bwd(g)>
    gg = torch.randn(2, rm.rows_dst, 5, generator=gen).numpy()
    lhs = float((got.astype(np.float64) * gg).sum())
    rhs = float((t0.reshape(2, -1, 5).numpy().astype(np.float64) * _apply(gg, rm.bwd_host)).sum())
    assert abs(lhs - rhs) <= 1e-9 * max(1.0, abs(lhs))


@pytest.mark.parametrize("dims,ws", GEOMS)
def test_voxel_map_equals_partition(dims, ws):
    _, g1 = _geoms(dims, ws)
    rm = G.rowmap_from_voxels(g1)
    x = torch.randn(2, 3, *dims, generator=torch.Generator().manual_seed(4))
    pads = R.pad_amounts(dims, ws)
    exp = R.partition_tokens(x, ws, R.effective_shift(dims, ws, tuple(w // 2 for w in ws)), pads)
    rows = x.permute(0, 2, 3, 4, 1).reshape(2, -1, 3).numpy()
    assert np.array_equal(_apply(rows, rm.fwd_host), exp.reshape(2, -1, 3).numpy())
    # every voxel is held by exactly one token slot
    assert np.array_equal(np.sort(rm.fwd_host[rm.fwd_host >= 0]), np.arange(int(np.prod(dims))))


def _ref_merge_gather(x, merge_last_dim):
    """The gather half of the reference PatchMerging.forward (down.py:21-47), restated with torch ops."""
    h, w, d = x.shape[2:]
    if h % 2 or w % 2 or d % 2:
        x = F.pad(x, tuple(reversed((0, h % 2, 0, w % 2, 0, d % 2))))
    if merge_last_dim:
        parts = [x[:, :, a::2, b::2, c::2] for a, b, c in
                 ((0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (1, 0, 1), (0, 1, 1), (1, 1, 1))]
    else:
        parts = [x[:, :, a::2, b::2, :] for a, b in ((0, 0), (1, 0), (0, 1), (1, 1))]
    t = torch.cat(parts, dim=1)
    return t.permute(0, 2, 3, 4, 1).reshape(x.shape[0], -1, t.shape[1]), tuple(t.shape[2:])


@pytest.mark.parametrize("dims,ws", GEOMS)
@pytest.mark.parametrize("mld", [True, False])
def test_merge_maps_equal_reference_gather(dims, ws, mld):
    _, g1 = _geoms(dims, ws)
    c = 3
    x = torch.randn(2, c, *dims, generator=torch.Generator().manual_seed(5))
    exp, mdims = _ref_merge_gather(x, mld)
    k = 8 if mld else 4
    rows = x.permute(0, 2, 3, 4, 1).reshape(2, -1, c).numpy()
    rm_v, md_v = G.rowmap_merge_from_voxels(tuple(dims), mld)
    assert md_v == mdims
    assert np.array_equal(_apply(rows, rm_v.fwd_host).reshape(2, -1, k * c), exp.numpy())
    # from block-output tokens: the output-side arrangement holds voxel out_map[slot]; cropped slots hold garbage
    rm_t, md_t = G.rowmap_merge(g1, mld)
    assert md_t == mdims
    out_map = g1.index_map_host(1).reshape(-1)
    tok = np.full((2, out_map.shape[0], c), 7.5, dtype=np.float32)
    ok = out_map >= 0
    tok[:, ok] = rows[:, out_map[ok]]
    assert np.array_equal(_apply(tok, rm_t.fwd_host).reshape(2, -1, k * c), exp.numpy())
