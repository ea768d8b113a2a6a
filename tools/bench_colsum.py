"""Graph-replay timing of pwa_colsum_rows against torch's reduction at the q|k|v gradient shapes of the bench step."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pwa_b200
from pwa_b200 import functional as PF

dev = "cuda:0"
for T, C in [(442368, 144), (55296, 288), (28672, 576), (442368, 48)]:
    xs = [torch.randn(T, C, device=dev).bfloat16() for _ in range(6)]
    for name, fn in [("pwa", PF.colsum_rows), ("torch", lambda t: t.sum(dim=0, dtype=torch.float32))]:
        fn(xs[0]); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for t in xs:
                o = fn(t)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 60
        print(f"T={T} C={C} {name}: {us:.1f} us  {T * C * 2 / us / 1e3:.0f} GB/s")
