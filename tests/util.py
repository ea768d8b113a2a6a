"""Shared helpers for the parity tests (golden loading, tolerance metric)."""
import glob
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names(prefix):
    return sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))


def load_npz(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


def load_block_case(name, dtype=torch.float64):
    """Returns (meta dict, state dict, x, p|None, go, out, grads dict) as torch tensors."""
    d = load_npz(name)
    m = [int(v) for v in d["meta"]]
    meta = dict(C=m[0], heads=m[1], ws=tuple(m[2:5]), dims=tuple(m[5:8]), shift=tuple(m[8:11]),
                I=m[11], E=m[12], B=m[13])
    sd = {}
    for k, v in d.items():
        if k.startswith("sd."):
            t = torch.from_numpy(v)
            sd[k[3:]] = t.to(dtype) if t.is_floating_point() else t
    x = torch.from_numpy(d["x"]).to(dtype)
    p = torch.from_numpy(d["p"]).to(dtype) if "p" in d else None
    go = torch.from_numpy(d["go"]).to(dtype)
    out = torch.from_numpy(d["out"]).to(torch.float64)
    grads = {k[5:]: torch.from_numpy(v).to(torch.float64) for k, v in d.items() if k.startswith("grad.")}
    return meta, sd, x, p, go, out, grads


def rel_linf(a, b):
    """max|a-b| / max|b|  -- the normalised L-inf metric of SURVEY.md §8(d) (atol = rtol*max|ref|)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    den = b.abs().max().item()
    return (a - b).abs().max().item() / (den if den > 0 else 1.0)
