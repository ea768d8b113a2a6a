import sys, torch
sys.path.insert(0, "/root/repo/tools")
dev = torch.device("cuda")
sys.path.insert(0, "/root/repo")
import pwa_b200
from pwa_b200 import functional as PF
def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
for T, Cout, Cin in ((28672, 576, 192), (13824, 192, 384), (3456, 384, 768), (55296, 96, 384)):
    NB = 4
    dys = [torch.randn(T, Cout, device=dev).bfloat16() for _ in range(NB)]
    xs = [torch.randn(T, Cin, device=dev).bfloat16() for _ in range(NB)]
    k = [0]
    def mm():
        k[0] = (k[0] + 1) % NB
        return torch.mm(dys[k[0]].t(), xs[k[0]], out_dtype=torch.float32)
    res = {"mm": round(timeit(mm), 1)}
    for S in (4, 8, 16, 32, 64):
        if T % S: continue
        def bm():
            k[0] = (k[0] + 1) % NB
            p = torch.bmm(dys[k[0]].view(S, T // S, Cout).transpose(1, 2), xs[k[0]].view(S, T // S, Cin), out_dtype=torch.float32)
            return PF.colsum_f32(p)
        res[f"bmm{S}"] = round(timeit(bm), 1)
    print(T, Cout, Cin, res, flush=True)
