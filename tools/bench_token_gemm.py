"""Timing of the fused token-GEMM kernel against the separate LayerNorm kernel + cuBLAS GEMM it replaces (graph replays)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pwa_b200
from pwa_b200 import functional as PF
sys.path.insert(0, ROOT)
from bench import _graph_time
dev = torch.device("cuda")
for (T, C) in ((4 * 432 * 256, 48), (4 * 54 * 256, 96), (4 * 28 * 256, 192)):
    xs = [torch.randn(T, C, device=dev).bfloat16() for _ in range(3)]
    rs = [torch.randn(T, C, device=dev).bfloat16() for _ in range(3)]
    g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    seed = torch.tensor([1, 2], dtype=torch.int32, device=dev)
    for Cout, mode in ((3 * C, "ln"), (C, "res_ln_bias"), (C, "drop_res_ln_bias")):
        W = torch.randn(Cout, C, device=dev).bfloat16()
        bias = torch.randn(Cout, device=dev).bfloat16() if "bias" in mode else None
        res = "res" in mode
        pd = 0.1 if "drop" in mode else 0.0
        with torch.no_grad():
            t_f = _graph_time(lambda i: PF._token_gemm_raw(xs[i % 3], rs[i % 3] if res else None, g, b, W, bias, res, True, True, 1e-6, pd, seed if pd else None), 10)

            def sep(i):
                x = xs[i % 3]
                if pd:
                    x = PF.seeded_dropout(x, pd, seed)
                if res:
                    s, z = PF.add_layer_norm(x, rs[i % 3], g, b, 1e-6)
                else:
                    z = PF.layer_norm(x, g, b, 1e-6)
                return torch.addmm(bias, z, W.t()) if bias is not None else torch.mm(z, W.t())
            t_s = _graph_time(sep, 10)
        nbytes = T * C * 2 * ((2 if res else 1) + (1 if res or pd else 0) + 1) + T * Cout * 2
        print(f"T={T} C={C} Cout={Cout} {mode:18s} fused {t_f:7.1f} us ({nbytes / t_f / 1e3:6.0f} GB/s)   separate {t_s:7.1f} us")
