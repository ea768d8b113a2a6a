"""torch.autograd glue around the C-ABI kernels.  PyTorch is plumbing here (device memory, streams,
autograd graph); all work is done by the sm_100a kernels in libpwa_b200.so.  CUDA tensors only."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from .geometry import Geometry

_DT = {torch.float32: _lib.PWA_F32, torch.bfloat16: _lib.PWA_BF16}


class KernelStats:
    """Launch accounting for bench.py: how many of OUR kernels were launched, and (optionally) CUDA-event
    timings around each C-ABI call on the launching stream.  Off by default (zero overhead)."""
    enabled = False
    timing = False
    launches = 0
    records = []          # (name, start_event, end_event, work) ; work = algorithmic FLOPs or bytes

    @classmethod
    def reset(cls, enabled=True, timing=False):
        cls.enabled, cls.timing, cls.launches, cls.records = enabled, timing, 0, []

    @classmethod
    def summary(cls):
        """name -> (calls, total_ms, total_work).  Call after torch.cuda.synchronize()."""
        out = {}
        for name, e0, e1, work in cls.records:
            c, ms, w = out.get(name, (0, 0.0, 0.0))
            out[name] = (c + 1, ms + e0.elapsed_time(e1), w + work)
        return out


class _timed:
    def __init__(self, name, n_kernels, work, t):
        self.name, self.n, self.work, self.t = name, n_kernels, work, t

    def __enter__(self):
        if KernelStats.enabled:
            KernelStats.launches += self.n
            if KernelStats.timing:
                self.e0 = torch.cuda.Event(enable_timing=True)
                self.e1 = torch.cuda.Event(enable_timing=True)
                self.e0.record(torch.cuda.current_stream(self.t.device))
        return self

    def __exit__(self, *exc):
        if KernelStats.enabled and KernelStats.timing:
            self.e1.record(torch.cuda.current_stream(self.t.device))
            KernelStats.records.append((self.name, self.e0, self.e1, self.work))
        return False


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype not in _DT:
        raise TypeError(f"pwa kernels support float32 and bfloat16, got {t.dtype}")
    return _DT[t.dtype]


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("pwa_b200 has no CPU path: tensors must live on a CUDA device (sm_100a)")


def _stream(t: torch.Tensor):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


# ------------------------------------------------------------------------------------------------
# (a) partition / reverse
# ------------------------------------------------------------------------------------------------
def _partition_raw(x: torch.Tensor, geom: Geometry, crop_lo: int, force_generic: bool = False,
                   force_word: bool = False) -> torch.Tensor:
    B, Cc = x.shape[:2]
    out = torch.empty((B, geom.P, geom.N, Cc), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device), _timed("partition", 1, 2.0 * out.numel() * out.element_size(), x):
        rc = _lib.lib.pwa_partition(_ptr(x), _ptr(out), B, Cc, geom.ref(), crop_lo | (2 if force_generic else 0) | (4 if force_word else 0),
                                    _dtype_code(x), _stream(x))
    _lib.check(rc, "pwa_partition")
    return out


def _reverse_raw(tok: torch.Tensor, geom: Geometry, crop_lo: int, force_generic: bool = False,
                 force_word: bool = False) -> torch.Tensor:
    B, _, _, Cc = tok.shape
    out = torch.empty((B, Cc, *geom.dims), dtype=tok.dtype, device=tok.device)
    with torch.cuda.device(tok.device), _timed("reverse", 1, 2.0 * tok.numel() * tok.element_size(), tok):
        rc = _lib.lib.pwa_reverse(_ptr(tok), _ptr(out), B, Cc, geom.ref(), crop_lo | (2 if force_generic else 0) | (4 if force_word else 0),
                                  _dtype_code(tok), _stream(tok))
    _lib.check(rc, "pwa_reverse")
    return out


class _Partition(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, geom):
        ctx.geom = geom
        return _partition_raw(x.contiguous(), geom, 0)

    @staticmethod
    def backward(ctx, g):
        # adjoint of a gather = scatter with the same (data-side) index map; padding slots drop out
        return _reverse_raw(g.contiguous(), ctx.geom, 0), None


class _Reverse(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tok, geom):
        ctx.geom = geom
        return _reverse_raw(tok.contiguous(), geom, 1)

    @staticmethod
    def backward(ctx, g):
        return _partition_raw(g.contiguous(), ctx.geom, 1), None


def partition_tokens(x: torch.Tensor, geom: Geometry) -> torch.Tensor:
    """[B,C,H,W,D] -> [B,P,N,C]: zero-pad, roll by -shift, strided window partition, channels last."""
    _require_cuda(x)
    if tuple(x.shape[2:]) != geom.dims:
        raise ValueError(f"feature map {tuple(x.shape[2:])} does not match geometry {geom.dims}")
    return _Partition.apply(x, geom)


def reverse_tokens(tok: torch.Tensor, geom: Geometry) -> torch.Tensor:
    """[B,P,N,C] -> [B,C,H,W,D]: window reverse, roll back, crop."""
    _require_cuda(tok)
    if tuple(tok.shape[1:3]) != (geom.P, geom.N):
        raise ValueError(f"token tensor {tuple(tok.shape)} does not match geometry P={geom.P} N={geom.N}")
    return _Reverse.apply(tok, geom)


# ------------------------------------------------------------------------------------------------
# (b)/(c) fused prompted window attention
# ------------------------------------------------------------------------------------------------
IMPL_AUTO, IMPL_F32, IMPL_TC = 0, 1, 2


def _shape_struct(B, P, Cc, heads, I, ws, scale, p_drop=0.0, seed=0, offset=0):
    s = _lib.PwaAttnShape()
    s.B, s.P, s.C, s.heads, s.I = B, P, Cc, heads, I
    s.ws[0], s.ws[1], s.ws[2] = ws
    s.scale, s.p_drop, s.seed, s.offset = float(scale), float(p_drop), int(seed), int(offset)
    return s


class _WindowAttention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, kp, vp, th, tw, td, tok, ids, heads, ws, scale, impl):
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        B, P, N, Cc = q.shape
        I = 0 if kp is None else kp.shape[1]
        if kp is not None:
            kp, vp, tok = kp.contiguous(), vp.contiguous(), tok.contiguous().float()
        th, tw, td = th.contiguous().float(), tw.contiguous().float(), td.contiguous().float()
        out = torch.empty_like(q)
        lse = torch.empty((B, P, heads, N), dtype=torch.float32, device=q.device)
        s = _shape_struct(B, P, Cc, heads, I, ws, scale)
        with torch.cuda.device(q.device), _timed("attn_fwd", 1, 4.0 * B * P * N * (N + I) * Cc, q):
            rc = _lib.lib.pwa_attn_fwd(_ptr(q), _ptr(k), _ptr(v), _ptr(kp), _ptr(vp), _ptr(th), _ptr(tw), _ptr(td),
                                       _ptr(tok), _ptr(ids), _ptr(out), _ptr(lse), C.byref(s), _dtype_code(q), impl,
                                       _stream(q))
        _lib.check(rc, "pwa_attn_fwd")
        ctx.save_for_backward(q, k, v, kp, vp, th, tw, td, tok, ids, out, lse)
        ctx.meta = (heads, tuple(ws), scale, impl, I)
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, kp, vp, th, tw, td, tok, ids, out, lse = ctx.saved_tensors
        heads, ws, scale, impl, I = ctx.meta
        B, P, N, Cc = q.shape
        dout = dout.contiguous()
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        f32 = dict(dtype=torch.float32, device=q.device)
        dkp = torch.empty((B, I, Cc), **f32) if I else None
        dvp = torch.empty((B, I, Cc), **f32) if I else None
        dth, dtw, dtd = torch.empty_like(th), torch.empty_like(tw), torch.empty_like(td)
        dtok = torch.empty_like(tok) if I else None
        delta = torch.empty_like(lse)
        s = _shape_struct(B, P, Cc, heads, I, ws, scale)
        with torch.cuda.device(q.device), _timed("attn_bwd", 2, 8.0 * B * P * N * (N + I) * Cc, q):
            rc = _lib.lib.pwa_attn_bwd(_ptr(q), _ptr(k), _ptr(v), _ptr(kp), _ptr(vp), _ptr(th), _ptr(tw), _ptr(td),
                                       _ptr(tok), _ptr(ids), _ptr(out), _ptr(lse), _ptr(dout), _ptr(dq), _ptr(dk),
                                       _ptr(dv), _ptr(dkp), _ptr(dvp), _ptr(dth), _ptr(dtw), _ptr(dtd), _ptr(dtok),
                                       _ptr(delta), C.byref(s), _dtype_code(q), impl, _stream(q))
        _lib.check(rc, "pwa_attn_bwd")
        if I:
            dkp, dvp = dkp.to(q.dtype), dvp.to(q.dtype)
        return dq, dk, dv, dkp, dvp, dth, dtw, dtd, dtok, None, None, None, None, None


def prompted_window_attention(q, k, v, kp, vp, th, tw, td, tok, ids, heads: int, ws: Sequence[int], scale: float,
                              impl: int = IMPL_AUTO) -> torch.Tensor:
    """q,k,v [B,P,N,C]; kp,vp [B,I,C] or None; th/tw/td [h,w,w] + tok [h,I] fp32 bias tables;
    ids uint8 [P,N] or None.  Returns [B,P,N,C].  See include/pwa.h: pwa_attn_fwd."""
    _require_cuda(q, k, v, kp, vp, th, tw, td, tok, ids)
    if q.shape[-1] % heads != 0:
        raise ValueError('WindowAttention: The dimension is not compatible with the number of heads!')
    return _WindowAttention.apply(q, k, v, kp, vp, th, tw, td, tok, ids, heads, tuple(ws), scale, impl)


# ------------------------------------------------------------------------------------------------
# LayerNorm over channels (+ fused residual add)
# ------------------------------------------------------------------------------------------------
class _AddLayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, res, gamma, beta, eps):
        x = x.contiguous()
        Cc = x.shape[-1]
        rows = x.numel() // Cc
        y = torch.empty_like(x)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        g32, b32 = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        s = None
        if res is not None:
            res = res.contiguous()
            s = torch.empty_like(x)
        nbytes = (2 if res is None else 4) * x.numel() * x.element_size()
        with torch.cuda.device(x.device), _timed("ln_fwd", 1, float(nbytes), x):
            rc = _lib.lib.pwa_ln_fwd(_ptr(x), _ptr(res), _ptr(g32), _ptr(b32), _ptr(s), _ptr(y), _ptr(mean), _ptr(rstd),
                                     rows, Cc, float(eps), _dtype_code(x), _stream(x))
        _lib.check(rc, "pwa_ln_fwd")
        normed = x if res is None else s
        ctx.save_for_backward(normed, g32, mean, rstd)
        ctx.has_res = res is not None
        ctx.param_dtypes = (gamma.dtype, beta.dtype)
        if res is None:
            return y
        return s, y

    @staticmethod
    def backward(ctx, *grads):
        normed, g32, mean, rstd = ctx.saved_tensors
        if ctx.has_res:
            ds, dy = grads
        else:
            (dy,), ds = grads, None
        Cc = normed.shape[-1]
        rows = normed.numel() // Cc
        dy = dy.contiguous() if dy is not None else torch.zeros_like(normed)
        ds = ds.contiguous() if ds is not None else None
        dx = torch.empty_like(normed)
        dg = torch.empty(Cc, dtype=torch.float32, device=normed.device)
        db = torch.empty(Cc, dtype=torch.float32, device=normed.device)
        nbytes = (3 if ds is None else 4) * normed.numel() * normed.element_size()
        with torch.cuda.device(normed.device), _timed("ln_bwd", 1, float(nbytes), normed):
            rc = _lib.lib.pwa_ln_bwd(_ptr(dy), _ptr(normed), _ptr(g32), _ptr(mean), _ptr(rstd), _ptr(ds), _ptr(dx), _ptr(dg),
                                     _ptr(db), rows, Cc, _dtype_code(normed), _stream(normed))
        _lib.check(rc, "pwa_ln_bwd")
        gd, bd = ctx.param_dtypes
        return dx, (dx if ctx.has_res else None), dg.to(gd), db.to(bd), None


def layer_norm(x, gamma, beta, eps: float = 1e-6):
    """LayerNorm over the last axis on the pwa kernel (C % 4 == 0, C <= 1024)."""
    _require_cuda(x, gamma, beta)
    return _AddLayerNorm.apply(x, None, gamma, beta, eps)


def add_layer_norm(x, res, gamma, beta, eps: float = 1e-6):
    """(s, y) with s = x + res and y = LayerNorm(s): the residual add of swin_block.py:222 fused into mlp_norm."""
    _require_cuda(x, res, gamma, beta)
    return _AddLayerNorm.apply(x, res, gamma, beta, eps)


def layer_norm_supported(C: int) -> bool:
    return C % 4 == 0 and C <= 1024
