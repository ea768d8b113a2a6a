"""Debug: weight-gradient GEMM dW = dy^T x for skinny outputs -- torch.mm (cuBLAS split-K) against a batched split over the
token axis (torch.bmm over S slices + a sum over the partials)."""
import sys, torch
dev = torch.device("cuda")

def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for _ in range(iters): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

for T, Cout, Cin in ((442368, 48, 48), (442368, 144, 48), (55296, 96, 96), (55296, 288, 96), (28672, 192, 192), (28672, 576, 192),
                     (55296, 96, 384), (6912, 192, 768)):
    NB = 4
    dys = [torch.randn(T, Cout, device=dev).bfloat16() for _ in range(NB)]
    xs = [torch.randn(T, Cin, device=dev).bfloat16() for _ in range(NB)]
    k = [0]
    def mm():
        k[0] = (k[0] + 1) % NB
        return torch.mm(dys[k[0]].t(), xs[k[0]], out_dtype=torch.float32)
    res = {"mm": round(timeit(mm), 1)}
    ref = torch.mm(dys[0].t().float(), xs[0].float())
    for S in (72, 144, 288, 576):
        if T % S: continue
        def bm():
            k[0] = (k[0] + 1) % NB
            p = torch.bmm(dys[k[0]].view(S, T // S, Cout).transpose(1, 2), xs[k[0]].view(S, T // S, Cin), out_dtype=torch.float32)
            return p.sum(0)
        try:
            res[f"bmm{S}"] = round(timeit(bm), 1)
            k[0] = -1
            err = (bm() - ref).abs().max().item() / ref.abs().max().item()
            res[f"err{S}"] = f"{err:.1e}"
        except Exception as e:
            res[f"bmm{S}"] = str(e)[:60]
    mb = T * (Cout + Cin) * 2 / 1e6
    print(T, Cout, Cin, f"{mb:.0f} MB", res, flush=True)
